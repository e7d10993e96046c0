#!/usr/bin/env python
"""profiles/ncu_traffic.json: DRAM bytes (read+write) per launch of each kernel, from an `ncu --set full` report.
usage: ncu_traffic.py <rep> <workload>"""
import csv, json, os, subprocess, sys
rep, wl = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = {"k_tc_forward": "tc_forward", "k_tc_grad_weight": "tc_grad_weight", "k_tc_grad_sample": "tc_grad_sample",
         "k_tc_grad_input": "tc_grad_input", "k_col2im_gather": "col2im_gather", "k_generic_forward": "generic_forward",
         "k_generic_bwd_data": "generic_bwd_data", "k_generic_bwd_weight": "generic_bwd_weight"}
out = {}
for r in rows[2:]:
    for k, n in names.items():
        if k + "<" in r[ki]:
            out[n] = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[wl] = out
json.dump(cur, open(path, "w"), indent=1)
print(out)
