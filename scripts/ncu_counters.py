#!/usr/bin/env python
"""ncu --set full reports -> profiles/kernel_counters.json: the per-launch DRAM / L2 counters bench.py quotes in `roofline`.
Each entry is keyed by kernel name and carries the workload shape it was captured on plus a hash of tepose_b200/csrc at capture
time; bench.py prints the counters only when name, shape and source hash all match the running build (else null).
usage: ncu_counters.py <shape-tag> report.ncu-rep [report.ncu-rep ...]     (shape-tag e.g. B32_T16_H2048_bf16 or smpl_16384_bf16)"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "kernel_counters.json")


def csrc_hash():
    h = hashlib.sha1()
    d = os.path.join(ROOT, "tepose_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".inl")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    db = json.load(open(OUT)) if os.path.isfile(OUT) else {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, units, r = rows[0], rows[1], rows[2]
        def val(name, scale_units=True):
            i = h.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "second": 1e6}.get(u, 1.0)
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].strip()
        e = {"shape": tag, "csrc_sha": csrc_hash(), "report": os.path.basename(rep),
             "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
             "l2_to_sm_read_bytes": 32.0 * val("lts__t_sectors_srcunit_tex_op_read.sum"),
             "duration_us_under_ncu": val("gpu__time_duration.sum"),
             "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "lts_throughput_pct": val("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
             "registers": val("launch__registers_per_thread")}
        db.setdefault(name, {})[tag] = e
        print(name, tag, {k: v for k, v in e.items() if k.endswith("bytes")})
    json.dump(db, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
