"""Loader: registers the hyphenated package directory `crnn-ocr-lite_b200/` under the importable name
`crnn_ocr_lite_b200` and re-exports it.  `import crnn_b200 as cb; cb.CRNN(...)`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crnn-ocr-lite_b200")
_NAME = "crnn_ocr_lite_b200"
if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
