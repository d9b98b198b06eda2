"""How much pipeline throughput does ONE extra (empty) kernel launch per frame cost? Adds k trivial
single-CTA launches (dodt_emit_detections into a dummy block) to every frame chain of the group
graphs and measures frames/s as bench.py does. usage: python tools/launch_cost_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops, shard, synth  # noqa: E402
from dodt_b200.frontend import FrontEnd, HostFrame  # noqa: E402

G, n_slots, steps = 8, 32, 1600
fe = FrontEnd()
slots = [fe.new_slot() for _ in range(n_slots)]
for i, s in enumerate(slots):
    HostFrame(fe).fill(synth.frame_inputs(2, i)).upload(s)
torch.cuda.synchronize()
dummy = shard.DetectionBlock(4, fe.cfg.avod_nms_size, fe.device)
streams = [torch.cuda.Stream() for _ in range(n_slots // G)]
main = torch.cuda.current_stream()
orig_post = FrontEnd._enqueue_post


def run(extra):
    def post(self, s, block, skip, rewrite=False):
        orig_post(self, s, block, skip, rewrite)
        for _ in range(extra):
            ops.emit_detections(s.prop_bev_boxes, s.final_scores, s.final_idx, s.n_final, dummy)
    FrontEnd._enqueue_post = post
    graphs = [fe.capture_group(slots[g * G:(g + 1) * G], slots[g * G - 1], None)[0] for g in range(n_slots // G)]

    def rr(n):
        for st in streams:
            st.wait_stream(main)
        for i in range(n // G):
            with torch.cuda.stream(streams[i % len(graphs)]):
                graphs[i % len(graphs)].replay()
        for st in streams:
            main.wait_stream(st)
    rr(3 * n_slots)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rr(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / steps


base = run(0)
print("0 extra launches: %.2f us/frame" % base)
for k in (4, 8, 16):
    t = run(k)
    print("%d extra launches per frame: %.2f us/frame (+%.2f us per launch)" % (k, t, (t - base) / k))
