"""oracle/ -- CPU restatement of the CGL-GAN hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the CPU baseline -- never as the thing shipped. The product
package (cgl-gan_b200/) must not import it.

Every function cites the reference file:line it restates (paths relative to /root/reference).
Parity pinning: the reference has no tests or golden vectors. The oracle is pinned
against outputs of the reference's own model classes and step bodies run in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
Unpinned parts (stated in DESIGN.md): fedlab's fedavg_aggregate / serialize_model (un-vendored
third-party dependency, restated from its published algorithm) and the LSGAN/MSE loss (no call site).
"""
