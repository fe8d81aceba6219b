"""kbo_b200 -- B200-native (sm_100a) implementation of kbo's k-bounded matching statistics hot path.

The compute lives in kbo_b200/libkbo_b200.so (hand-written CUDA behind the C ABI of
include/kbo_b200.h); `kbo_b200.api` mirrors the reference crate's public functions over it.
"""
