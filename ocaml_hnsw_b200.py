"""Import shim: the package directory is `ocaml-hnsw_b200/` (the repo layout names it after the
reference); a hyphen cannot be imported, so `import ocaml_hnsw_b200` loads that directory as a
package under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ocaml-hnsw_b200")
_spec = importlib.util.spec_from_file_location(
    "ocaml_hnsw_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ocaml_hnsw_b200"] = _mod
_spec.loader.exec_module(_mod)
