#!/usr/bin/env python
"""Large-N Pulsatile fixtures from the UNTOUCHED reference, started from the open vessel at rest.

    make -C oracle ref && python tests/golden/make_golden_pulsatile_open.py

The reference's own start is a vessel closed at the inlet whatever N, which the reference cannot advance for N >= 128
(DESIGN.md 3.5); the large-N tests and bench.py therefore start from pulsatile_cases.open_vessel_at_rest.  This script
hands that very state to the untouched header (oracle/_ref/ref_pulsatile state=FILE: the reference re-derives the mask,
Fobj and the border lists from the wall positions with its own functions) and records what the REFERENCE makes of it:
SHA-256 of P, Ux, Uy, yr1, yr2, the node mask and both lattice buffers at the dump steps (the arrays are tens of MB), the wall
positions themselves, and a few scalars that show the walls moved and fresh nodes appeared.
-> tests/golden/pulsatile_open_sha256.json, checked against the oracle by tests/test_pulsatile_oracle.py.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "ref_pulsatile")
CASES = {"open_N128_m6": (128, 6.0, [1, 100, 600]), "open_N256_m6": (256, 6.0, [300])}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    cases = entry.load_package().pulsatile_cases
    out = {}
    for name, (N, margin, dumps) in CASES.items():
        nx, ny = 1 + 10 * (N - 2), N
        ne = nx * ny
        st = cases.open_vessel_at_rest(N, margin=margin)
        with tempfile.TemporaryDirectory() as td:
            sf, df = os.path.join(td, "state.bin"), os.path.join(td, "dump.bin")
            with open(sf, "wb") as f:
                for k in ("lattice", "P", "Ux", "Uy", "yr1", "yr2"):
                    f.write(np.ascontiguousarray(st[k], dtype=np.float64).tobytes())
            log = subprocess.check_output([EXE, "N=%d" % N, "steps=%d" % max(dumps), "state=" + sf, "out=" + df,
                                           "dump_at=" + ",".join(map(str, dumps))]).decode().strip()
            print(name, log)
            raw = np.fromfile(df, dtype=np.uint8)
        rec = {"N": N, "margin": margin, "dumps": dumps, "steps": {}}
        off = 0
        for d in dumps:
            e = {}
            arr = {}
            for nm, n in (("P", ne), ("Ux", ne), ("Uy", ne), ("yr1", nx), ("yr2", nx)):
                arr[nm] = raw[off:off + 8 * n].view(np.float64)
                e[nm] = sha(arr[nm])
                off += 8 * n
            flag = raw[off:off + ne]
            e["flag"] = sha(flag)
            off += ne
            e["lattice"] = sha(raw[off:off + 8 * 18 * ne])
            off += 8 * 18 * ne
            e["parity"] = int(raw[off:off + 4].view(np.int32)[0])
            off += 4
            # readable evidence that this is not a frozen state: wall travel, open nodes, peak velocity
            e["yr1_min_max"] = [float(arr["yr1"].min()), float(arr["yr1"].max())]
            e["yr2_min_max"] = [float(arr["yr2"].min()), float(arr["yr2"].max())]
            e["bulk_nodes"] = int((flag == 1).sum())
            e["Ux_max"] = float(np.abs(arr["Ux"]).max())
            rec["steps"][str(d)] = e
        assert off == raw.size
        assert sha(st["flag"]) is not None
        rec["initial_bulk_nodes"] = int((st["flag"] == 1).sum())
        out[name] = rec
    json.dump(out, open(os.path.join(HERE, "pulsatile_open_sha256.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
