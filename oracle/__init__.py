"""TEST INFRASTRUCTURE — CPU oracles for the SNCH-LBVH hot path.

* ``liboracle.so``  : plain-C restatement of the reference algorithm (``snch_oracle.c``), every function cites the
  reference file:line it follows.  Pinned against the reference's own code through ``_ref/libsnch_ref_cpu.so``
  (tests/test_oracle_pinning.py) and the committed golden vectors under tests/golden/.
* ``_ref/*.so``     : the UNMODIFIED reference headers compiled where they lie under /root/reference
  (see oracle/Makefile); git-ignored build outputs.

Only tests/, bench.py's baseline legs and __graft_entry__.smoke() may import this package.  The product
(``snch-lbvh_b200``) never does and has no CPU fallback.
"""
from .loader import RefScene, RefScene2, OracleScene, OracleScene2, FcpwScene, ref_available, oracle_lib, build_oracle, host_libm, restated_libm  # noqa: F401
