#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
(time timeout 900 python -m pytest tests -m gpu -q) 2>&1 | grep "passed\|failed\|^FAILED\|real"
timeout 300 python bench.py --steps 10 --no-cpu-baseline --sub-records fleet --sub-scans 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'fleet',round(d['sub_records']['sequences_fleet']['value'],1))"
