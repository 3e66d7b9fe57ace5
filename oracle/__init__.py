"""CPU oracle for the VAN-GAN hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (torch fp32/fp64 + an independent numpy/scipy
version of the stencil ops), the arithmetic of the reference's volumetric
train-step / sliding-window path (psweens/VAN-GAN).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; the product package never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and
its own implementation (TensorFlow 2.10.1 + tensorflow_addons 0.20.0 + keras
2.10) cannot be imported in this image (not installed, no network).  The
third-party semantics below are therefore recalled, not verified against TF:

* tfa InstanceNormalization: eps=1e-3, biased variance, y = gamma*xhat+beta.
* TF 'SAME' padding for even kernels puts the extra pad AFTER (k4 s1: 1 / 2).
* Conv3D k1 s2 'same' has no padding and samples even indices.
* tf.pad REFLECT excludes the border voxel (index -1 -> 1).
* Keras BinaryCrossentropy(from_logits=False): clip to [1e-7, 1-1e-7], then
  -(y*log(p+1e-7) + (1-y)*log(1-p+1e-7)), mean over the channel axis.
* UpSampling3D(2): nearest repeat.  MaxPool3D 'same': out-of-range excluded.
* max-pool / min / max gradient ties: single winner (first in scan order) for
  pooling; reduce_min/reduce_max split evenly.  For soft_skel the input
  gradient is independent of WHICH tied element wins as long as exactly one
  does (ties after an erosion always carry the same source voxel's value).
* Keras OptimizerV2 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t), eps=1e-7 outside
  the sqrt; order = all-reduce SUM -> per-variable clip_by_norm -> apply.

What pins the oracle instead: an independent numpy/scipy restatement of the
stencil/loss ops (bitwise-equal checks), fp64 finite-difference checks of the
backward passes, and analytic known-answer cases (tests/test_oracle_*.py);
small golden vectors produced by this oracle are frozen under tests/golden/.
"""
