"""Child process (HB_EMUL_TENSOR=1): tools/lbs_forms_time.measure - the verification bench.py's `roofline_candidates` children
run on the B200 - executed on the CPU emulation at a small size, so that its logic is tested before it meets hardware."""
import json
import sys

root, lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + '/tests/host/emul')
sys.path.insert(0, root + '/tools')
import cpu_backend  # noqa: E402

cpu_backend.install(lib)
import lbs_forms_time as F  # noqa: E402

forms = [tuple(int(x) for x in f.split(',')) for f in sys.argv[3].split(';')]
recs = F.measure(forms, B=3, T=43, reps=1, slab=512, device='cpu', peak_gbs=6490.5)
print(json.dumps({'recs': recs}))
