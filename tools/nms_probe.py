"""Phase timing of nms_round (reads the diagnostic block of the NMS workspace)."""
import os, sys, struct
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops
    from oracle import synth_ref as synth
from dodt_b200._lib import load
anch = synth.car_anchors()
rng = np.random.default_rng(0)
kept = np.sort(rng.choice(len(anch), 60000, replace=False))
_, boxes, scores = synth.rpn_proposals(2, 0, anch[kept])
def run(b, s, max_out, thr, label, max_windows=0):
    b = torch.from_numpy(b).cuda(); s = torch.from_numpy(s).cuda()
    n = b.shape[0]
    ws = torch.zeros(ops.nms_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        keep, nk = ops.nms(b, s, max_out, thr, workspace=ws, max_windows=max_windows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); keep, nk = ops.nms(b, s, max_out, thr, workspace=ws, max_windows=max_windows); e1.record()
    torch.cuda.synchronize()
    off = int(load().dodt_nms_state_offset(n))
    raw = ws[off:off + 64].cpu().numpy().tobytes()
    ints = struct.unpack("4i", raw[:16]); t = struct.unpack("6Q", raw[16:64])
    print("  [%s] n_kept=%d done=%d sweeps=%d | launch->tiles %.1f us, mask load %.1f, solve %.1f, emit %.1f" %
          (label, ints[0], ints[1], ints[3], (t[1]-t[0])/1e3, (t[2]-t[1])/1e3, (t[3]-t[2])/1e3, (t[4]-t[3])/1e3))
    return keep, nk, e0.elapsed_time(e1) * 1e3
k, nk, us = run(boxes, scores, 1024, 0.8, "rpn")
print("rpn 60k: %.1f us" % us, nk.cpu().tolist())
top = k[:int(nk[0])].cpu().numpy()
fs = rng.permutation(np.linspace(0.01, 0.99, len(top))).astype(np.float32)
k2, nk2, us2 = run(boxes[top], fs, 100, 0.01, "final")
print("final 1024: %.1f us" % us2, nk2.cpu().tolist())
