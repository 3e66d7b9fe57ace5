"""epoch_sweep -- same signature and behaviour as the reference's post_training.py:4-39: for every saved epoch in
range(start, end + 1, step) restore the checkpoint from <output_dir>/checkpoints, make <output_dir>/Epoch_Sampling/e{i} and run
the sliding-window mapping of every file under `test_path` into it (stride 50, padFactor 0.1, file prefix e{i}_VG_)."""
import os


def epoch_sweep(args, vangan_model, plotter, test_path='', start=100, end=200, step=2, segmentation=True):
    out = {}
    for i in range(start, end + 1, step):
        print(i)
        vangan_model.load_checkpoint(epoch=i, newpath=args.output_dir + '/checkpoints')
        filepath = args.output_dir + '/Epoch_Sampling/'
        folder = os.path.join(filepath, 'e{idx}'.format(idx=i))
        if not os.path.isdir(folder):
            os.makedirs(folder)
        # the reference indexes the os.listdir() list with the file NAME (post_training.py:35-36, a TypeError); what it means is:
        testfiles = [os.path.join(test_path, file) for file in sorted(os.listdir(test_path))]
        filename = 'e{idx}_VG_'.format(idx=i)
        out[i] = plotter.run_mapping(vangan_model, testfiles, args.INPUT_IMG_SIZE, filetext=filename, segmentation=segmentation,
                                     stride=(50, 50, 50), filepath=folder, padFactor=0.1)
    return out
