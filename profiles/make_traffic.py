"""DRAM traffic per op from `ncu --set full` captures (what bench.py reports as roofline.traffic).
usage: python profiles/make_traffic.py out.json rep1.ncu-rep [rep2.ncu-rep ...]
The captures are those of scripts/collect_evidence.sh (scripts/roi_probe.py ncu, scripts/fpn_bwd_probe.py)."""
import csv
import json
import subprocess
import sys

OPS = {  # op of bench.py -> kernels launched for it (substring match on the kernel name)
    "roi_fuse_fwd": ["roi_prep_kernel", "roi_fuse_fwd_ring"],
    "roi_fuse_bwd": ["roi_bin_kernel", "roi_bwd_pull_tma"],
    "fpn_bwd": ["fpn_bwd_fused_tma", "gather_bwd_up_cl"],
}


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    res = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        rd = wr = None
        for k, v in d.items():
            if k == "dram__bytes_read.sum":
                rd = float(v.replace(",", ""))
            if k == "dram__bytes_write.sum":
                wr = float(v.replace(",", ""))
        unit = {h: u for h, u in zip(hdr, rows[1])}["dram__bytes_read.sum"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        res.setdefault(name, []).append(((rd or 0) * scale, (wr or 0) * scale, float(d.get("gpu__time_duration.sum", "0").replace(",", ""))))
    return res


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    ks = {}
    for rep in reps:
        ks.update(kernels(rep))
    doc = {"source": "ncu --set full --clock-control none captures on a B200 (" + ", ".join(reps) + "); bytes = "
                     "dram__bytes_read.sum + dram__bytes_write.sum per launch, last captured launch of each kernel, "
                     "bench.py's workload (2 images x 512 RoIs, C = 256, fp32)"}
    for op, subs in OPS.items():
        tot, parts = 0.0, {}
        for sub in subs:
            for name, launches in ks.items():
                if sub in name:
                    rd, wr, _ = launches[-1]
                    parts[sub] = {"read": rd, "write": wr}
                    tot += rd + wr
        if parts:
            doc[op] = {"bytes": tot, "kernels": parts}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
