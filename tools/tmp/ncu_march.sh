set -e
cd /root/repo
cat > /tmp/run_win10.py <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ripcurrents_b200 import Context, synth
dev = torch.device('cuda', 0)
cc = Context(0)
B = 8; ww, hh = 1920, 1080
frames = torch.from_numpy(np.stack(synth.clip(ww, hh, B + 1, seed=1))).to(dev)
cc.flow_configure_batch(ww, hh, 0.5, 2, 20, 3, 15, 1.2, 256, B)
cc.flow_push_batch(frames.data_ptr(), count=1)
for _ in range(2):
    cc.flow_push_batch(frames.data_ptr() + ww * hh, count=B)
cc.synchronize(); cc.close()
PY
python /tmp/run_win10.py
ncu --set full --clock-control none --import-source on -k regex:flow_march -s 8 -c 2 -o gpurun_out/march_full -f python /tmp/run_win10.py > gpurun_out/ncu_march.log 2>&1
ls -la gpurun_out/
