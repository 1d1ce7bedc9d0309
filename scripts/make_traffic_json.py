#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --set full` capture of the queue kernel (runs without a GPU).

    python scripts/make_traffic_json.py gpurun_out/r02d_persist.ncu-rep C4 20 > profiles/traffic.json

bench.py accepts the figure only when the captured kernel name, grid and block size are the ones
it launches (lig_pick_kernel_info), so a stale capture cannot vouch for a different kernel."""
import csv
import json
import re
import subprocess
import sys


def main(path, workload, steps_per_launch):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]

    def get(k):
        i = hdr.index(k)
        v = float(r[i])
        u = units[i]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    name = re.sub(r"^void\s+", "", r[hdr.index("Kernel Name")])
    short = re.match(r"[A-Za-z_0-9:]+", name).group(0).split("::")[-1]
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    out = {
        "_comment": "DRAM traffic of the dominant kernel from ncu --set full --clock-control none (one cold, "
                    "serialised launch): dram__bytes_read.sum + dram__bytes_write.sum.  Reads = the descriptors exactly "
                    "once (+ the compact tables once per CTA from L2); the last tens of MB of picks are still dirty in "
                    "the 126 MB L2 when the single profiled launch ends, so writes come out below the algorithmic 8 B/decision.",
        workload: {
            "capture": {"kernel": short, "full_name": name, "grid": int(float(r[hdr.index("launch__grid_size")])),
                        "threads": int(float(r[hdr.index("launch__block_size")])),
                        "steps_per_launch": int(steps_per_launch), "file": path.split("/")[-1],
                        "duration_us": float(r[hdr.index("gpu__time_duration.sum")])},
            "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
            "bytes_per_launch": int(rd + wr), "bytes_per_step": int((rd + wr) / int(steps_per_launch)),
        },
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
