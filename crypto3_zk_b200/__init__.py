"""crypto3_zk_b200 - B200-native (sm_100a) kernels for crypto3-zk's two data-parallel proving hot
paths (NTT/LDE/FRI-commit and MSM) behind the reference's call-site API.

Layout:  csrc/   CUDA kernels + the C ABI (include/zkb200.h) -> libzkb200.so
         host/   C++ host templates mirroring the reference's math::/algebra:: entities
         api.py  Python host layer over the same ABI (tests, bench)
"""
from . import capi  # noqa: F401
from .api import Context, MerkleTree, MsmBases, field_generator, msm_combine, unity_root  # noqa: F401
from .fields import CURVE_BY_NAME, CURVES, FIELD_BY_NAME, FIELDS  # noqa: F401
