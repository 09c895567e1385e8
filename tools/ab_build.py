"""Build variant libraries for tools/ab.sh:  python tools/ab_build.py name=DEF1,DEF2 name2= ..."""
import os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshoptimizer_b200 import build
vdir = os.path.join(build.LIB_DIR, "variants")
shutil.rmtree(vdir, ignore_errors=True)
for arg in sys.argv[1:]:
    name, _, defs = arg.partition("=")
    print(build.build_variant(name, [d for d in defs.split(",") if d]))
