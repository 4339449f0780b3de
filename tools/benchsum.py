#!/usr/bin/env python
"""Print the headline numbers of bench.py JSON lines: python tools/benchsum.py file.log [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]),
              {k: round(v, 4) for k, v in d["roofline"]["per_class_ms_per_md_step"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
