"""Ad-hoc GPU check used during bring-up: engine vs oracle on seeded corpora with
first-mismatch diagnostics.  usage: python tools/gpu_debug.py [model kind n] ..."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import corpus  # noqa: E402
from oracle import oracle_py  # noqa: E402
from sentencepiece_b200 import Engine  # noqa: E402


def check(model, kind, n, seed=99, tune=None):
    mb = open(os.path.join(ROOT, "tests", "golden", "models", model + ".model"), "rb").read()
    g = corpus.CorpusGen()
    buf, offs = g.fill(kind, seed, n)
    eng = Engine(mb)
    if tune:
        eng.set_tuning(*tune)
    t = time.time()
    ids, ido = eng.encode_packed(buf, offs)
    dt = time.time() - t
    info = eng.info()
    om = oracle_py.OracleModel(mb)
    oids, oido = om.encode_batch(buf, offs)
    ok = np.array_equal(ido, oido) and np.array_equal(ids, oids)
    print(f"{model}/{kind} n={n} tune={tune}: {'OK' if ok else 'MISMATCH'} ids={len(ids)} oracle_ids={len(oids)} "
          f"wall={dt*1e3:.1f}ms main_kernel={info.last_main_kernel_ms:.3f}ms all={info.last_kernel_ms:.3f}ms "
          f"deferred={info.last_deferred} launches={info.last_kernel_launches}", flush=True)
    if not ok:
        raw = buf.tobytes()
        bad = 0
        for i in range(n):
            a = ids[int(ido[i]):int(ido[i + 1])]
            b = oids[int(oido[i]):int(oido[i + 1])]
            if not np.array_equal(a, b):
                bad += 1
                if bad <= 3:
                    s = raw[int(offs[i]):int(offs[i + 1])]
                    print("  sentence", i, repr(s))
                    print("   engine:", a.tolist())
                    print("   oracle:", b.tolist())
                    print("   norm  :", om.normalize(s)[0])
        print("  mismatching sentences:", bad)
    eng.close()
    return ok


if __name__ == "__main__":
    args = sys.argv[1:]
    if args:
        ok = True
        for i in range(0, len(args), 3):
            ok &= check(args[i], args[i + 1], int(args[i + 2]))
        sys.exit(0 if ok else 1)
    ok = True
    for m, k, n in [("uni32k", "en", 20000), ("uni32k", "mixed", 5000), ("mix_bf8k", "mixed", 5000),
                    ("botchan8k", "en", 5000), ("bpe32k", "en", 5000), ("mix_bpe4k", "mixed", 5000)]:
        try:
            ok &= check(m, k, n)
        except Exception as e:  # keep going: bring-up wants all the information it can get
            print(f"{m}/{k}: EXCEPTION {e!r}", flush=True)
            ok = False
    sys.exit(0 if ok else 1)
