"""Regenerates tests/golden/* from the REFERENCE ITSELF (oracle/_ref/libydref.so,
i.e. /root/reference's task_dispatcher.cc compiled verbatim).  Run in the dev
container:  python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from yadcc_b200 import TaskDispatcher  # noqa: E402
from yadcc_b200 import streams as S  # noqa: E402

REF = ROOT / "oracle" / "_ref" / "libydref.so"
NAMES = ["cfg1", "cfg2-mod-small", "cfg2-random-small", "cfg3-small", "cfg3-mod-small", "cfg-self-small"] + [f"fuzz-{i}" for i in range(40)]
BIG = ["cfg2-mod", "cfg2-random", "cfg-self"]  # full BASELINE sizes; ~2-4 s each on the reference
# BASELINE configs[2] at full size (1 M x 4 k, three solve / free-half / tick rounds: 2.7 M decisions) and
# configs[4]'s pool and distributions on the first 1 M requests of its queue: minutes each on the reference
HUGE = ["cfg3", "cfg5-1m"]


def main():
    """No arguments: everything.  With stream names: only those, merged into the existing file."""
    path = Path(__file__).parent / "digests.json"
    only = sys.argv[1:]
    if only:
        out = json.loads(path.read_text())
    else:
        out = {"generator": "oracle/_ref/libydref.so (reference compiled verbatim)", "streams": {}}
    for name in only or NAMES + BIG + HUGE:
        d = TaskDispatcher(str(REF))
        r = S.Replayer(d)
        tr = r.run(S.named_stream(name, d))
        out["streams"][name] = {"sha256": S.trace_digest(tr), "decisions": r.decisions, "granted": r.granted}
        if name == "cfg1":
            g = tr[0]
            np.savez_compressed(Path(__file__).parent / "cfg1_reference.npz", status=g["status"],
                                task_id=g["task_id"], servant_index=g["servant_index"])
        d.close()
        print(name, out["streams"][name])
    path.write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
