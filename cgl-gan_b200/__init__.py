"""cgl-gan_b200 -- B200-native engine for the CGL-GAN simulated-client hot path.

Layout:
  csrc/         CUDA kernels + the C ABI (include/cgl_b200.h), built into lib/libcgl_b200.so
  abi.py        ctypes binding of the C ABI (fails loudly when the library is missing)
  layout.py     packed parameter rows <-> reference state_dicts
  models.py     Generator / Discriminator classes with the reference's interfaces
  generators.py server-side generators stacked over edge servers
  engine.py     ClientBank: packed per-client discriminators driven through the C ABI
  partition.py  bit-exact dataset partitioning / topology / client selection (host Python)
  sim.py        host loops (CGLGAN, CAPGAN, Mix-G, MDGAN, ACGAN, FLGAN, FeGAN) with the reference's knobs
  dist.py       one-process-per-GPU sharding and the NCCL aggregation hook
"""
__version__ = "0.1.0"
