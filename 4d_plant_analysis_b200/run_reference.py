"""``python -m local_icp_b200.run_reference <reference script> [args...]``: runs an unmodified driver
of the reference (run_registration_pipeline.py, ...) with the GPU hot path installed.  A
``sys.meta_path`` hook patches the two reference modules right after they are imported, whatever
``sys.path`` games the driver plays (run_registration_pipeline.py:8,14)."""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import runpy
import sys

_TARGETS = {"registration.registration_icp", "p2p_matching_in_organ.p2p_matching_seg_icp"}


class _PatchAfterImport(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname not in _TARGETS:
            return None
        sys.meta_path.remove(self)
        try:
            spec = importlib.util.find_spec(fullname)
        finally:
            sys.meta_path.insert(0, self)
        if spec is None or spec.loader is None:
            return None
        loader = spec.loader
        orig_exec = loader.exec_module

        def exec_module(module, _orig=orig_exec):
            _orig(module)
            from .install import install
            install()

        loader.exec_module = exec_module
        return spec


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m local_icp_b200.run_reference <script.py> [args...]")
    script = os.path.abspath(argv[0])
    sys.meta_path.insert(0, _PatchAfterImport())
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
