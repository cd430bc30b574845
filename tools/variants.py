"""Experiment helper: build scan-kernel variants here (CPU box), time them on the GPU box.
   python tools/variants.py build   -> gr_air_modes_b200/variants/<name>.so
   python tools/variants.py run     -> (on the GPU) swaps each variant in and runs tests/tools/prof_time.py
   python tools/variants.py run-decode [log2n ...] -> same with tests/tools/prof_decode.py --check (decoder experiments)"""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "gr_air_modes_b200", "variants")
VARIANTS = {"hi128": ["AMB_SCAN_REGS_HI=128"], "hi120": ["AMB_SCAN_REGS_HI=120"], "hi112": ["AMB_SCAN_REGS_HI=112"],
            }
if sys.argv[1] == "build":
    from gr_air_modes_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    for name, d in VARIANTS.items():
        print(build.build_native(force=True, defines=d, out_path=os.path.join(VDIR, name + ".so")))
else:
    lib = os.path.join(ROOT, "gr_air_modes_b200", "libairmodes_b200.so")
    keep = lib + ".keep"
    shutil.copy(lib, keep)
    try:
        for name in sorted(os.listdir(VDIR)):
            if not name.endswith(".so"): continue
            shutil.copy(os.path.join(VDIR, name), lib)
            if sys.argv[1] == "run-decode":
                if not name.startswith("pair"): continue
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "prof_decode.py"), "--check"] + sys.argv[2:], capture_output=True, text=True)
                print(name, out.stdout.strip() if out.stdout.strip() else out.stderr[-300:])
                continue
            if sys.argv[1] == "run-holes":
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_holes.py")] + sys.argv[2:], capture_output=True, text=True)
                print(name, "\n".join(out.stdout.strip().splitlines()[-2:]) if out.stdout.strip() else out.stderr[-300:])
                continue
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "prof_time.py")] + sys.argv[2:], capture_output=True, text=True)
            print(name, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
    finally:
        shutil.copy(keep, lib); os.remove(keep)
