"""oracle/ — TEST INFRASTRUCTURE, not product code.

A restatement of the reference's algorithm for the inter-frame interpolation
path, used only as the checker: by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / `--impl reference` leg.  Nothing under
flood_uav_video_segmentation_b200/ imports it.

Where the arithmetic lives.  The reference (pure Python) delegates every
number on this path to a third-party dependency that is not vendored under
/root/reference: torch (pinned ==1.11.0 in Pipfile:7; this image has 2.11.0)
and numpy (pinned 1.24.3; this image has 2.3).  The restatement therefore
comes in two layers:

  flow_oracle.py / metric_oracle.py
      the reference's *call sequence* (flow/model.py, flow/base.py,
      util/util.py, base/foundation.py) re-expressed over the same torch /
      numpy calls, each function citing the file:line it follows.  Run on
      torch-CPU it is the CPU baseline; run on torch-CUDA on the same B200 it
      is the bit-exact authority for label maps (torch-CPU and torch-CUDA
      grid_sample differ at the 1e-5 level).
  c/fuvs_oracle.c
      a plain-C restatement of the *published ATen CUDA algorithm* those
      calls execute (GridSampler.cuh, UpSample.cuh, the elementwise mul/add,
      max-with-indices, histc), compiled with -ffp-contract=off and explicit
      fmaf() where nvcc fuses.  It lets the CPU-only test tier check golden
      vectors without a GPU.

Pinning.  The reference holds no tests, golden vectors or fixtures for this
path ("parity unpinned" upstream).  The oracle is pinned
instead against outputs of the reference itself, executed in the authoring
container by importing /root/reference (make_golden.py, committed), stored as
small fixtures under tests/golden/.  tests/test_oracle_golden.py re-checks the
oracle against those fixtures on every run, and
tests/test_oracle_vs_reference.py re-imports the live reference when
/root/reference is present.
"""
