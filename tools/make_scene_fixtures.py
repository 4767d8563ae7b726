"""Regenerate tests/scenes/*.json from the reference's example inputs (run in the build container,
where /root/reference exists).  The scenes named by BASELINE.json's configs are re-serialised as
compact JSON (all numbers float64, exactly the values the reference's YAML/JSON loaders produce)
so that tests, smoke() and bench.py can run on the GPU box, where /root/reference is absent."""
import json
import sys
from pathlib import Path

import yaml

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples")
OUT = Path(__file__).resolve().parents[1] / "tests" / "scenes"
OUT.mkdir(parents=True, exist_ok=True)
for src in sorted(REF.iterdir()):
    if src.suffix not in (".yaml", ".json"):
        continue
    data = yaml.safe_load(src.read_text()) if src.suffix == ".yaml" else json.loads(src.read_text())
    dst = OUT / (src.stem + ".json")
    dst.write_text(json.dumps(data, separators=(",", ":"), sort_keys=True) + "\n")
    print(dst.name, dst.stat().st_size)
