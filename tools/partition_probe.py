"""Spatial partitioning probe: the S4 launches of the frame pipeline confined to `s4_sms` SMs through a
CUDA green context (torch.cuda.green_contexts), the per-frame chains everywhere else. The group graphs
are split in a pre part (S1, S2, S3a, S5a) and a post part (S3b, S5b); S4 is launched eagerly on the
green context's stream (8 pairs per launch, CTA count 2 x s4_sms) and tied in with events.
usage: python tools/partition_probe.py [s4_sms (0 = no partition, plain low-priority stream)] [sweeps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dodt_b200 import ops, synth  # noqa: E402
from dodt_b200.frontend import FrontEnd, FrontEndConfig, HostFrame  # noqa: E402

s4_sms = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n_slots, GROUP = 32, 8
cfg = FrontEndConfig()
fe = FrontEnd(cfg)
c = cfg
slots = [fe.new_slot() for _ in range(n_slots)]
for i, s in enumerate(slots):
    HostFrame(fe).fill(synth.frame_inputs(2, i)).upload(s)
torch.cuda.synchronize()
n_groups = n_slots // GROUP
streams = [torch.cuda.Stream() for _ in range(n_groups)]
main = torch.cuda.current_stream()
PRE_SKIP, POST_SKIP = ("S4", "S3b", "S5b"), ("S1", "S2", "S3a", "S5a", "S4")
pre, post = [], []
for g in range(n_groups):
    grp, prev = slots[g * GROUP:(g + 1) * GROUP], slots[g * GROUP - 1]
    pre.append(fe.capture_group(grp, prev, None, PRE_SKIP)[0])
    post.append(fe.capture_group(grp, prev, None, POST_SKIP)[0])
if s4_sms > 0:
    from torch.cuda.green_contexts import GreenContext
    gctx = GreenContext.create(s4_sms, torch.cuda.current_device())
    s4_stream = gctx.Stream()
    max_ctas = 2 * s4_sms
else:
    gctx = None
    s4_stream = torch.cuda.Stream(priority=0)
    max_ctas = 0
print("S4 stream:", s4_stream, "max_ctas", max_ctas, flush=True)
ev_s4 = [torch.cuda.Event() for _ in range(n_groups)]
ev_post = [torch.cuda.Event() for _ in range(n_groups)]


def s4(g):
    grp, prev = slots[g * GROUP:(g + 1) * GROUP], slots[g * GROUP - 1]
    ops.correlation_stream([prev.bev_feat] + [s.bev_feat for s in grp], 1, c.corr_max_displacement, 1,
                           c.corr_stride_2, c.corr_padding, outs=[s.corr for s in grp], max_ctas=max_ctas)


def rr(n):
    for st in streams:
        st.wait_stream(main)
    s4_stream.wait_stream(main)
    for i in range(n * n_groups):
        g = i % n_groups
        with torch.cuda.stream(s4_stream):
            s4_stream.wait_event(ev_post[g])          # the previous sweep's crops of these corr maps are done
            s4(g)
            ev_s4[g].record(s4_stream)
        with torch.cuda.stream(streams[g]):
            pre[g].replay()
            streams[g].wait_event(ev_s4[g])
            post[g].replay()
            ev_post[g].record(streams[g])
    for st in streams:
        main.wait_stream(st)
    main.wait_stream(s4_stream)


for e, st in zip(ev_post, streams):
    e.record(st)
rr(4)
torch.cuda.synchronize()
# reference result of one slot from the regular single-graph path, for a sanity check of the split form
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
rr(sweeps)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) * 1e3 / (sweeps * n_slots)
print("s4_sms %d: %.1f us/frame (%.0f frames/s)" % (s4_sms, us, 1e6 / us), flush=True)
chk = slots[5].corr.clone()
want = ops.correlation(slots[4].bev_feat, slots[5].bev_feat, 1, c.corr_max_displacement, 1, c.corr_stride_2, c.corr_padding)
print("corr of slot 5 equals the pairwise call:", bool(torch.equal(chk, want)), "final n:", int(slots[5].n_final[0].item()))
