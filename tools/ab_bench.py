"""A/B harness: run bench.py's main() against an alternative build of the library (tools/build_variant.sh).
    python tools/ab_bench.py dl-dkd_b200/variants/libdkd_b200_X.so [bench args...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = os.path.abspath(sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
import dkd_b200._lib as L  # noqa: E402
L.LIB_PATH = lib
import bench  # noqa: E402
bench.main()
