"""ofdm_b200 -- B200-native OFDM baseband engine (CUDA kernels behind a C ABI).

Layout: csrc/ (sm_100a kernels + the C ABI), engine.py (ctypes binding of include/ofdm_engine.h),
api.py (host-side mirror of the reference crate's encode/decode/Header/Analysis API).
"""
from .engine import (  # noqa: F401
    Config, ChannelParams, Engine, EngineError, RxResult, load_library,
    MOD_BPSK, MOD_QPSK, MOD_QAM64, SYNC_REFERENCE, SYNC_SCHMIDL_COX, CFO_REFERENCE, CFO_ANGLE_OF_SUM,
    PHASE_REFERENCE, PHASE_ANGLE_OF_SUM, OK, TOO_SHORT, NO_SYNC, BAD_HEADER, NEG_OFFSET,
)
from .api import (  # noqa: F401
    Analysis, DecodeError, Header, ModulationScheme, bytes_to_sig, decode, decode_batch, encode, encode_batch,
    sig_to_bytes,
)
