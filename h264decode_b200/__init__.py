"""h264decode_b200 -- B200-native (sm_100a) data-parallel front end of an H.264 decoder.

The product is the C-ABI shared library libh264b200.so (include/h264b200.h; sources in csrc/).  `capi` is a thin
ctypes binding of that ABI used by tests/ and bench.py; it contains no compute and there is no CPU fallback:
importing `h264decode_b200.capi` without the built library raises ImportError, and creating a Context without a
CUDA device raises H264BError(NO_DEVICE).  (`h264decode_b200.build` builds the library and is importable without it.)
"""


def __getattr__(name):  # lazy, so that `python -m h264decode_b200.build` works before the library exists
    if name in ("capi", "Context", "H264BError"):
        import importlib
        capi = importlib.import_module(".capi", __name__)
        return capi if name == "capi" else getattr(capi, name)
    raise AttributeError(name)
