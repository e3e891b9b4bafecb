#!/usr/bin/env python3
"""Writes tests/golden/go_tables.json: the reference's Go table literals (rangeTabLPS, stateTransxTab, MNVars,
CodedblockPatternMN) as read by tests/ref_model.py's own reader from /root/reference.  The GPU box has no
/root/reference; the second model loads this snapshot there.  tests/test_ref_model.py re-checks the snapshot against
the Go files whenever they are present."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_model  # noqa: E402

t = ref_model.read_go_tables()
with open(ref_model.SNAPSHOT, "w") as f:
    json.dump(ref_model._to_json(t), f, sort_keys=True, separators=(",", ":"))
print("wrote", ref_model.SNAPSHOT, {k: len(v) for k, v in t.items()})
