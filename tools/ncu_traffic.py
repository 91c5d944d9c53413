"""Writes profiles/ncu_traffic.json from `ncu --page raw --csv` dumps of the coder kernels: DRAM bytes (read + write) per
launch, the hash of the kernel sources they were measured with and the commit.  bench.py reports the numbers as
roofline.traffic as long as the kernel sources still hash to the same value (null afterwards).

usage: python tools/ncu_traffic.py <workload>:<frames>=<raw.csv> [...]      e.g. cfg2:128=gpurun_out/prof_r2a_raw.csv"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_sha  # noqa: E402

entries = {}
for spec in sys.argv[1:]:
    key, _, path = spec.partition("=")
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    body = rows[rows.index(hdr) + 2 :]
    i_name, i_r, i_w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[rows.index(hdr) + 1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in body:
        if len(r) != len(hdr):
            continue
        kernel = "k_encode_tiled" if "k_encode_tiled" in r[i_name] else "k_decode_tiled" if "k_decode_tiled" in r[i_name] else None
        if kernel:
            total = float(r[i_r].replace(",", "")) * scale.get(units[i_r], 1) + float(r[i_w].replace(",", "")) * scale.get(units[i_w], 1)
            entries[f"{kernel}:{key}"] = int(total)
git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {
    "_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, tools/r2_prof.sh); kernels_sha = bench.kernel_sources_sha() "
                "of the sources the capture ran; bench.py reports roofline.traffic from here only while that hash still matches",
    "kernels_sha": kernel_sources_sha(), "git": git, "entries": entries,
}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=2)
print(json.dumps(out, indent=2))
