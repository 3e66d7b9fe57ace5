"""epoch_sweep — same signature as the reference's post_training.py:4-39: for each saved epoch, load the
checkpoint and run the sliding-window mapping over the test set."""
import os


def epoch_sweep(args, vangan_model, plotter, test_path='', start=100, end=200, step=10, segmentation=True):
    test_files = sorted(os.path.join(test_path, f) for f in os.listdir(test_path) if f.endswith(".npy"))
    out = {}
    for epoch in range(start, end + 1, step):
        ckpt = os.path.join(args.output_dir, "checkpoints", "checkpoint_e%d.npz" % epoch)
        if not os.path.exists(ckpt):
            print("Error: Checkpoint not found!", ckpt)
            continue
        vangan_model.load_checkpoint(ckpt)
        filepath = os.path.join(args.output_dir, "e%d" % epoch)
        os.makedirs(filepath, exist_ok=True)
        out[epoch] = plotter.run_mapping(vangan_model, test_files, args.INPUT_IMG_SIZE, segmentation=segmentation,
                                         stride=(50, 50, 50), padFactor=0.1, filetext="e%d_" % epoch, filepath=filepath)
    return out
