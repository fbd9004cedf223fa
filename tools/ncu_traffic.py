"""profiles/r2_traffic.json from an `ncu --set full` capture of the hot kernels of one bench config:
DRAM bytes (read + write) per launch, averaged per kernel class, keyed like bench.py's per-kernel
list, stamped with the digest of the CUDA sources the capture was taken on (bench.py quotes it as
roofline.traffic only while the sources are unchanged).
usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep [config] [out.json]"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(rep, config="voc32", out=None):
    out = out or os.path.join(os.path.dirname(__file__), "..", "profiles", "r2_traffic.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, ik, ig = (hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"),
                      hdr.index("launch__grid_size"))
    acc = {}
    # the larger-grid launches of a kernel class are the bilateral (d = 5) lattice, the smaller the Gaussian (d = 2)
    recs = []
    for d in data:
        name = d[ik]
        cls = "blur" if "blur_kernel" in name else "splat" if ("splat_coop" in name or "splat_fast" in name) else \
            "slice" if "slice_softmax" in name else None
        if cls is None:
            continue
        recs.append((cls, to_bytes(d[ir], units[ir]) + to_bytes(d[iw], units[iw]), int(float(d[ig].replace(",", "")))))
    # blur: the larger grid is the bilateral (d = 5) lattice; the splats are persistent (same grid for both
    # lattices) and are told apart by the bytes they move (the bilateral one writes 5x more lattice rows)
    grids = sorted({g for c, _, g in recs if c == "blur"})
    for c, b, g in recs:
        if c == "blur":
            acc.setdefault("blur_kernel d=%d" % (5 if (len(grids) == 1 or g == grids[-1]) else 2), []).append(b)
    sb = sorted(b for c, b, _ in recs if c == "splat")
    if sb:
        cut = (sb[0] + sb[-1]) / 2 if sb[-1] > 1.2 * sb[0] else 0
        for c, b, g in recs:
            if c == "splat":
                acc.setdefault("splat_kernel d=%d" % (5 if b >= cut else 2), []).append(b)
    for c, b, g in recs:
        if c == "slice":
            acc.setdefault("slice_softmax_kernel (fused 2 terms)", []).append(b)
    js = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches) from one "
                      "`ncu --set full --clock-control none` capture; bench.py copies the dominant kernel's entry into "
                      "roofline.traffic while csrc_sha256_16 matches the sources",
          "config": config, "csrc_sha256_16": bench.csrc_digest(), "source": os.path.basename(rep),
          "kernels": {k: int(sum(v) / len(v)) for k, v in acc.items()},
          "launches": {k: len(v) for k, v in acc.items()}}
    json.dump(js, open(out, "w"), indent=1)
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
