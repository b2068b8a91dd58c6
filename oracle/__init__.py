"""CPU oracle for the patch-graph BA + correlation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / CPU baseline.  The product
path (``cdv-slam_b200/``) never imports this package and fails loudly when its CUDA
library is missing.

Contents
  se3_np.py        numpy SE3 helpers restating cdvslam/fastba/ba_cuda.cu:36-174
  ba_oracle.py     float64 numpy restatement of cuda_ba() (cdvslam/fastba/ba_cuda.cu:232-611,
                   block_e.cu:43-300): the parity target of the drop-in
  ba_torch_port.py torch/CPU restatement of the reference's torch path (cdvslam/ba.py:86-185 +
                   cdvslam/projective_ops.py:53-113): the ``cpu_baseline`` ("port") that bench.py times
  corr_oracle.py   numpy restatement of altcorr corr / patchify (correlation_kernel.cu:17-47, 83-136,
                   193-233; correlation.py:51-71) and of the reproject kernel (ba_cuda.cu:408-458)
  neighbors_oracle.py  restatement of cuda_ba.neighbors (ba.cpp:59-97)

Parity pinning: the reference ships no golden vectors for this path.  The oracle is
pinned instead against outputs of the reference's own ``cdvslam/ba.py`` + ``projective_ops.py`` imported
verbatim in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), see
DESIGN.md "Oracle".
"""
