"""Every evidence file DESIGN.md / README.md / INTEGRATION.md / BASELINE.md cite under profiles/ is committed, and the ncu traffic table
bench.py reads names existing sources."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("doc", ["DESIGN.md", "README.md", "INTEGRATION.md", "BASELINE.md"])
def test_cited_profiles_exist(doc):
    text = open(os.path.join(ROOT, doc)).read()
    missing = []
    for m in re.finditer(r"profiles/([A-Za-z0-9_\-\.\*\{\},]+)", text):
        name = m.group(1).rstrip(".,);:")
        if any(c in name for c in "*{"):          # patterns such as r02s_*.json: at least one match
            pat = re.sub(r"\{[^}]*\}", "*", name)
            import glob
            if not glob.glob(os.path.join(ROOT, "profiles", pat)):
                missing.append(name)
        elif not os.path.exists(os.path.join(ROOT, "profiles", name)):
            missing.append(name)
    assert not missing, "%s cites profiles/ files that are not committed: %s" % (doc, missing)


def test_ncu_traffic_sources_exist():
    d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for k, v in d.items():
        if k.startswith("_"):
            continue
        assert v["bytes_per_launch"] > 0
        assert os.path.exists(os.path.join(ROOT, v["source"])), (k, v["source"])
