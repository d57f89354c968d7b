"""CPU oracle for the GDN hot path -- TEST INFRASTRUCTURE ONLY.

Plain fp32 PyTorch/numpy restatements of the reference algorithms (each function
cites the reference file:line it follows).  Only tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() may import this package; the
product path (gdn_pytorch_b200/) never does and fails loudly without its CUDA
library.

Pinning: the reference has no tests or golden vectors of its own,
so the oracle is pinned against outputs of the unmodified reference imported in
the dev container (oracle/gen_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py, and live against /root/reference when present).
"""
