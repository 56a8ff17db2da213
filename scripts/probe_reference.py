#!/usr/bin/env python
"""Is the real implementation of the hot path (cellpose==4.0.8, what /root/reference imports) available here?
Prints one JSON line; exit code 0 either way.  Run it first on any new box:

    python scripts/probe_reference.py

When it reports "available": true, `python -m pytest tests/test_real_cellpose.py` pins the oracle against it and
`bench.py --impl reference` / `cpu_baseline` time it (kind "reference") instead of the oracle port."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import real  # noqa: E402

r = real.find()
if r is None:
    print(json.dumps({"available": False, "tried": real.find.tried,
                      "pinned_version": "cellpose==4.0.8 (uv.lock:352-353), fastremap==1.17.7, fill-voids==2.1.1"}))
else:
    print(json.dumps({"available": True, "version": r["version"], "where": r["where"]}))
