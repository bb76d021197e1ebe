"""pingpongpro_b200 -- B200-native (sm_100a) implementation of PingPongPro's data-parallel core.

The product is native: `csrc/` (CUDA kernels behind the C ABI of include/ppp_b200.h -> lib/libppp_b200.so) and
`host/` (SAM/BAM decode, annotation, writers, the drop-in `pingpongpro` command line -> lib/libppp_host.so,
bin/pingpongpro).  The Python modules are thin ctypes bindings for the tests and bench.py:

    api   -- ppp_b200.h (contexts, stack build, scan, scoring, transposons); no CPU fallback
    host  -- ppp_host.h (file decode, annotation reader, synthetic reads)
    dist  -- collective hooks for one process per GPU (torch.distributed or the library's NCCL binding)
    _build -- explicit g++ / nvcc command lines that build everything in-tree
"""
