"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration of the kernels bench.py's
roofline entries name from `ncu --set full` reports, and write profiles/r2_ncu_traffic.json (read by bench.py).
usage: python scripts/ncu_traffic.py <tag>=<report.ncu-rep>:<kernel substring>:<batch> ..."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    res = {}
    for spec in sys.argv[1:]:
        tag, rest = spec.split("=", 1)
        rep, sub, batch = rest.split(":")
        hdr, units, rows = load(rep)
        ki = hdr.index("Kernel Name")
        picked = [r for r in rows if sub in r[ki]]
        if not picked:
            print("no kernel matching", sub, "in", rep)
            continue
        r = picked[-1]

        def val(name):
            i = hdr.index(name)
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        dur_i = hdr.index("gpu__time_duration.sum")
        dur = float(r[dur_i].replace(",", ""))
        du = units[dur_i]
        dur_us = dur / 1000.0 if du in ("ns", "nsecond") else (dur if du in ("us", "usecond") else dur * 1000.0)
        res[tag] = {"kernel": r[ki][:120], "batch": int(batch), "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                    "duration_us_under_ncu": dur_us, "report": Path(rep).name}
        print(tag, res[tag])
    out = ROOT / "profiles" / "r2_ncu_traffic.json"
    old = json.loads(out.read_text()) if out.exists() else {}
    old.update(res)
    out.write_text(json.dumps(old, indent=1))


if __name__ == "__main__":
    main()
