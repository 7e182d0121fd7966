#!/usr/bin/env python
"""profiles/traffic.json from an ncu summary (tools/ncu_summary.py output): DRAM bytes per launch of every kernel of the
step, keyed the way bench.py names them.  Usage: make_traffic.py profiles/r02_ncu_summary.jsonl > profiles/traffic.json"""
import json, re, sys
KEYS = [("k_count_part", "count_part"), ("k_tab_apply_marked", "tab_apply"), ("k_enum_count", "enum_count"), ("k_enum_lin", "enum_lin"),
        ("k_part_bounds", "count_bounds"), ("k_rp_pass", "partition_pass"), ("k_rp_hist", "partition_hist"), ("k_rp_scan", "partition_scan"), ("k_seg_scan", "enum_scan"), ("DeviceRadixSortOnesweep", "partition_sort_pass_portion"),
        ("k_ec_lookup", "ec_lookup"), ("k_ec_cov", "ec_cov"), ("k_ec_setup", "ec_setup"), ("k_ec_rescue", "ec_rescue"), ("k_ec_ext", "ec_ext"), ("k_ec_search", "correct"),
        ("k_ec_merge", "ec_merge"), ("k_trim", "trim")]
def num(s):
    v, u = s.split()[0], (s.split() + [""])[1]
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "Tbyte": 1e12, "byte": 1.0, "ms": 1.0, "us": 1e-3, "": 1.0}.get(u, 1.0)
out = {}
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    for pat, key in KEYS:
        if pat in d.get("kernel", "") and "dram_read" in d:
            rd, wr = num(d["dram_read"]), num(d["dram_write"])
            if key in out and out[key]["dram_bytes_per_launch"] >= rd + wr:
                continue  # keep the largest launch of a kernel (the full window)
            out[key] = {"kernel": re.sub(r"\(.*", "", d["kernel"]).replace("void ", ""), "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                        "duration_under_ncu_ms": num(d["duration"]), "source": sys.argv[1] + " (ncu, one launch = one window of the full-size step)"}
            break
json.dump(out, sys.stdout, indent=1)
