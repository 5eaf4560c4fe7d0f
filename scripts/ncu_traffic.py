"""profiles/ncu_traffic.json: dram bytes (read + write) per launch of each kernel class, averaged over the
launches captured in the given `ncu --set full` reports.  usage: ncu_traffic.py class=report.ncu-rep[:name-filter] ..."""
import csv, io, json, os, subprocess, sys
out = {}
for arg in sys.argv[1:]:
    cls, rest = arg.split("=", 1)
    rep, _, flt = rest.partition(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik, it = (hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"),
                      hdr.index("gpu__time_duration.sum"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n, names = 0.0, 0, set()
    for r in rows[2:]:
        if flt and flt not in r[ik]:
            continue
        tot += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        n += 1
        names.add(r[ik].split("(")[0][:60])
    out[cls] = {"bytes_per_launch": round(tot / max(n, 1)), "launches": n, "kernels": sorted(names),
                "source": os.path.basename(rep)}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
