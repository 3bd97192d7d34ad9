"""Launcher: run one of the reference's scripts, unmodified, on the compat overlay.

    python compat/run.py /path/to/nerfmeshes/src/mesh_nerf.py --log-checkpoint <run dir> --res 256 ...

Python puts a script's own directory first on sys.path, which would resolve `models` / `nerf` to the reference's
packages; this launcher puts compat/ (and the repo root) ahead of it and then executes the script as __main__.
"""
import os
import runpy
import sys

if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    here = os.path.dirname(os.path.abspath(__file__))
    script = os.path.abspath(sys.argv[1])
    sys.path[:0] = [here, os.path.dirname(here), os.path.dirname(script)]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")
