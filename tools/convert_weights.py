"""Checkpoint / ONNX -> FPB2W weight file (what `tools/cvt_onnx2trt.bash` is to the reference).

  python tools/convert_weights.py refiner_hwc.onnx refiner.fpb2w --kind refiner
  python tools/convert_weights.py scorer.pth scorer.fpb2w --kind scorer
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foundationpose_cpp_b200 import weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("src")
ap.add_argument("dst")
ap.add_argument("--kind", required=True, choices=["refiner", "scorer"])
a = ap.parse_args()
weights.import_weights(a.src, a.dst, weights.KIND_REFINER if a.kind == "refiner" else weights.KIND_SCORER)
print(f"wrote {a.dst} ({os.path.getsize(a.dst) / 1e6:.1f} MB)")
