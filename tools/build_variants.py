"""Build kernel variants into flac_raster_b200/lib/var_<name>.so: tools/build_variants.py name=-DX=1,-DY=2 ..."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flac_raster_b200 import _native as nat
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = nat.LIB_PATH.parent / f"var_{name}.so"
    cmd = ["nvcc", *nat.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-o", str(out), str(nat.CSRC / "flacraster_b200.cu")]
    procs.append((name, subprocess.Popen(cmd)))
for name, p in procs:
    print(name, "rc", p.wait())
