"""TEST INFRASTRUCTURE ONLY -- runs code under the reference's OWN runtime, in the container that has /root/reference.

Loads, from where they lie, the reference's `core/base.py` (ModuleBase / ModuleManager, with the one line that does not
import on Python 3.12 patched IN MEMORY: the mutable dataclass default at core/base.py:521), its own cffi binding
`core/bindings/camera_message_framework.py` (bound to the transport compiled unmodified into oracle/_ref), its
`core/capture_source.py` and `capture_sources/image_directory.py`.  The packages the reference expects from the rest of
the CUAUV tree and that are absent here are stubbed: `auv_python_helpers` (library lookup), `auvlog.client` (logging),
`shm` (the vehicle's variable store).  Nothing of the reference is copied into the repository; nothing here can run on
the GPU box (no /root/reference there), which is why tests/test_real_runtime.py is a CPU-container test.
"""
import importlib
import os
import sys
import types

REF = os.environ.get("BV_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMF_LIB = os.path.join(ROOT, "oracle", "_ref", "libcamera_message_framework.so")


def available():
    return os.path.isfile(os.path.join(REF, "core", "base.py")) and os.path.isfile(CMF_LIB)


class _Log:
    """auvlog.client.log: attribute access makes child loggers, calling one logs (core/base.py:645, capture_source.py:24)."""

    def __init__(self, path="", sink=None):
        self._path, self._sink = path, sink if sink is not None else []

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return _Log(self._path + "." + name, self._sink)

    def __call__(self, msg, copy_to_stdout=False):
        self._sink.append((self._path, str(msg)))


_installed = None


def install():
    """Registers the stubs and the `vision` package alias; returns the loaded reference modules."""
    global _installed
    if _installed is not None:
        return _installed
    helpers = types.ModuleType("auv_python_helpers")
    helpers.get_library_path = lambda name: os.path.join(ROOT, "oracle", "_ref", name)
    helpers.load_library = lambda name: __import__("ctypes").CDLL(helpers.get_library_path(name))
    sys.modules["auv_python_helpers"] = helpers
    auvlog = types.ModuleType("auvlog")
    client = types.ModuleType("auvlog.client")
    client.Logger = _Log
    client.log = _Log()
    auvlog.client = client
    sys.modules["auvlog"] = auvlog
    sys.modules["auvlog.client"] = client
    sys.modules.setdefault("shm", types.ModuleType("shm"))
    vision = types.ModuleType("vision")
    vision.__path__ = [REF]                      # `vision.core.base` == /root/reference/core/base.py
    sys.modules["vision"] = vision
    cmf = importlib.import_module("vision.core.bindings.camera_message_framework")
    tuners = importlib.import_module("vision.core.tuners")
    capture = importlib.import_module("vision.core.capture_source")
    # core/base.py with line 521 patched in memory: `_acquisition_times: Deque[int] = deque(maxlen=30)`
    path = os.path.join(REF, "core", "base.py")
    src = open(path).read()
    bad = "_acquisition_times: Deque[int] = deque(maxlen=30)"
    assert src.count(bad) == 1, "core/base.py changed: the Python 3.12 patch no longer applies"
    src = src.replace(bad, "_acquisition_times: Deque[int] = __import__('dataclasses').field("
                           "default_factory=lambda: deque(maxlen=30))")
    base = types.ModuleType("vision.core.base")
    base.__file__ = path
    sys.modules["vision.core.base"] = base
    exec(compile(src, path, "exec"), base.__dict__)
    image_directory = importlib.import_module("vision.capture_sources.image_directory")
    _installed = types.SimpleNamespace(base=base, cmf=cmf, tuners=tuners, capture=capture, image_directory=image_directory,
                                       log=client.log)
    return _installed
