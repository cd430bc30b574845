"""profiles/scan_traffic.json from the ncu --set full captures of the scan kernel (gpurun_out/scan_{4,10,20}msps.ncu-rep):
dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by bench config, together with the sha256 of the kernel
source the captures were taken of - bench.py only reports `roofline.traffic` while that is the source that is built."""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gr_air_modes_b200", "csrc", "amb_kernels.cu")
sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
out = {}
for key, rep in (("c1", "scan_4msps"), ("c2", "scan_10msps"), ("c3", "scan_20msps")):
    path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    def val(name):
        i = h.index(name); x = float(v[i].replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[i]]
    out[key + "_2p28"] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                          "kernel": v[h.index("Kernel Name")], "kernels_sha256": sha,
                          "source": "ncu --set full --clock-control none -k regex:amb_scan -s 1 -c 1 python tools/prof_run.py 28 <rate> 2 (profiles/r2_%s_ncu.txt)" % rep}
json.dump(out, open(os.path.join(ROOT, "profiles", "scan_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
