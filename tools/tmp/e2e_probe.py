import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ripcurrents_b200 import Context, synth, capi
W, H, B = 1920, 1080, 32
dev = torch.device('cuda', 0)
frames = synth.clip(W, H, B + 1, seed=0)
order = list(range(B + 1)) + list(range(B - 1, 0, -1))
seq = np.stack([frames[i] for i in order])
h_seq = torch.from_numpy(seq).pin_memory(); d_seq = h_seq.to(dev)
h_masks = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
d_masks = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]
res = [(capi.FrameResult * B)() for _ in range(2)]
NB = W * H * B
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
def run(mode, steps=30):
    ctx = Context(0); ctx.set_stream(stream.cuda_stream)
    ctx.flow_configure_batch(W, H, 0.5, 2, 3, 2, 15, 1.2, 0, B); ctx.hist_reset(); ctx.window_configure(W, H, 10)
    st = {'s': 0}
    def step():
        s = st['s']
        src = (h_seq if mode in ('full', 'h2d_only') else d_seq).data_ptr() + (s & 1) * NB
        if mode == 'full': m = h_masks[s & 1].data_ptr(); r = res[s & 1]
        elif mode == 'h2d_only': m = None; r = None
        elif mode == 'd2h_only': m = h_masks[s & 1].data_ptr(); r = res[s & 1]
        elif mode == 'dev_masks': m = d_masks[s & 1].data_ptr(); r = None
        else: m = None; r = None
        ctx.process_frames(src, 31 + s * B, m, count=B, submit_only=True, results=r, want_results=r is not None)
        st['s'] = s + 1
    for _ in range(4): step()
    ctx.wait(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): step()
    ctx.wait(); e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.close()
    return ms
for mode in ('device', 'dev_masks', 'h2d_only', 'd2h_only', 'full', 'device'):
    ms = run(mode)
    print('%-10s %.3f ms/step  %.0f pairs/s' % (mode, ms, B / ms * 1e3), flush=True)
