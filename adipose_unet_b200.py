"""Import shim: registers the on-disk package directory ``adipose_tissue-unet_b200/``
(the name the build contract fixes; a hyphen is not importable) as the Python
package ``adipose_unet_b200``.  ``import adipose_unet_b200`` then behaves like a
normal package import, submodules included."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adipose_tissue-unet_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
