"""ctypes binding of libsniper_b200.so (the C-ABI declared in include/sniper_b200.h).

The product path has no CPU fallback: if the library is missing, loading raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsniper_b200.so")

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_long, "f": ctypes.c_float, "z": ctypes.c_size_t,
      "d": ctypes.c_double, "u": ctypes.c_uint}

# name -> (restype code, argument codes); order and meaning documented in include/sniper_b200.h
SIGNATURES = {
    "sniper_last_error": ("s", ""),
    "sniper_abi_version": ("i", ""),
    "sniper_multi_proposal_target_workspace_bytes": ("z", "iiii"),
    "sniper_generate_anchors": ("i", "ipipip"),
    "sniper_proposal_decode": ("i", "pppp" "iiiii" "pipi" "iii" "ppp" "p"),
    "sniper_multi_proposal_target_fwd": ("i", "ppppp" "iiiiiii" "pipi" "f" "iii" "pppp" "pp" "pz" "p"),
    "sniper_deform_psroi_fwd": ("i", "ppp" "iiii" "f" "iiiii" "f" "iii" "ppp" "p"),
    "sniper_deform_psroi_bwd": ("i", "pppp" "iiii" "f" "iiiii" "f" "iii" "pp" "p"),
    "sniper_psroi_fwd": ("i", "pp" "iiii" "f" "iiii" "pp" "p"),
    "sniper_psroi_bwd": ("i", "pp" "iiii" "f" "iiii" "p" "p"),
    "sniper_gemm_nt": ("i", "plplpl" "iiii" "ppp" "l" "iii" "p"),
    "sniper_conv2d_nhwc": ("i", "piiii" "pii" "pp" "iii" "pl" "iiiii" "i" "ppp" "l" "iii" "p"),
    "sniper_conv2d_wgrad_nhwc": ("i", "pp" "iiiii" "i" "pp" "iii" "p" "ii" "p"),
}

_lib = None


def lib():
    """Returns the loaded C-ABI library; raises (loudly) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "sniper_b200: %s is missing -- run `python -m sniper_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = ctypes.c_char_p if res == "s" else _T[res]
            fn.argtypes = [_T[c] for c in args]
        _lib = L
    return _lib


class SniperError(RuntimeError):
    """Raised when a C-ABI entry point returns non-zero (MXNetError analogue)."""


def check(rc):
    if rc != 0:
        raise SniperError(lib().sniper_last_error().decode())
