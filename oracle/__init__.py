"""oracle/ — CPU restatement of the reference's algorithm for the try-on hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under the product package (`shineon-virtual-tryon_b200/`) may import
this; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do,
and there only as the checker / the timed CPU baseline.

Every function cites the reference file:line it follows (paths relative to the reference checkout).
Arithmetic that the reference delegates to PyTorch ATen (conv2d, instance_norm, batch_norm, bmm,
grid_sample, upsample, gelu, softmax, linear, torch.inverse; pinned pytorch=1.6.0 in sams-pt1.6.yaml:112,
run here on torch 2.11 CPU) is delegated to the same ATen functions through torch.nn.functional — the
restatement is of the reference's own code (module graphs, index conventions, CUDA kernels), not of ATen.

Parity pin: the reference ships no golden vectors, so the oracle is
pinned against outputs of the reference's own classes, imported read-only in the build container by
`oracle/make_golden.py` and stored under tests/golden/.  The three CUDA-only reference ops
(Resample2d / Correlation / ChannelNorm) cannot be executed without a GPU in that container; their
restatements are pinned only through the reference-built `oracle/_ref` extensions when those compile
(see oracle/build_ref.py) — otherwise "parity unpinned" applies to those three ops.
"""
