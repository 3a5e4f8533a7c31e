"""Time simon_host_simulate on the full C3 objects (10,000 nodes, 120,500 pods): phases as reported by the library.
    python tools/native_c3.py [REPS]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "open-simulator_b200"))
sys.path.insert(0, ROOT)
from simon_b200 import native_host, synth  # noqa: E402

cluster, apps = synth.make_c3()
req = native_host.request_json(cluster, apps)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    nat = native_host.simulate_native(None, None, req=req)
    print(json.dumps({"call_s": round(native_host.simulate_native.last_call_s, 4), "decode_s": round(native_host.simulate_native.last_decode_s, 4),
                      **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in nat["timing"].items()}}))
