#!/usr/bin/env python
"""Extract dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from `ncu --set full`
captures (gpurun_out/*.ncu-rep) into profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic.
usage: python profiles/update_traffic.py workload=report.ncu-rep[:kernel-regex[:max-launches]] ..."""
import csv, io, json, os, re, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_traffic.json")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def rows_of(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        d["_units"] = dict(zip(hdr, units))
        yield d


def val(d, key):
    return float(d[key].replace(",", "")) * UNIT[d["_units"][key]]


def main():
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in sys.argv[1:]:
        wl, spec = arg.split("=", 1)
        rep, _, rest = spec.partition(":")
        rx, _, mx = rest.partition(":")
        mx = int(mx) if mx else 10 ** 9
        tot_r = tot_w = 0.0
        names, grid = [], None
        for d in rows_of(rep):
            if rx and not re.search(rx, d["Kernel Name"]):
                continue
            if len(names) >= mx:
                break
            tot_r += val(d, "dram__bytes_read.sum")
            tot_w += val(d, "dram__bytes_write.sum")
            names.append(d["Kernel Name"].split("(")[0])
            grid = d.get("Grid Size")
        # "summary" (the committed text summary of the report under profiles/, added by hand) survives a re-extraction of the same report
        keep = {"summary": table[wl]["summary"]} if table.get(wl, {}).get("report") == os.path.basename(rep) and "summary" in table[wl] else {}
        table[wl] = {"kernels": names, "dram_bytes_read": tot_r, "dram_bytes_write": tot_w, "traffic": tot_r + tot_w,
                     "grid": grid, "report": os.path.basename(rep), **keep}
        print(wl, names, "%.3f GB read, %.3f GB written" % (tot_r / 1e9, tot_w / 1e9))
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
