import sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ripcurrents_b200 import Context
dev = torch.device('cuda', 0)
c = Context(0)
w, h = 1920, 1080
fl = torch.randn((h, w, 2), device=dev)
img = torch.zeros((h, w, 3), dtype=torch.uint8, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):
    c._chk(c.lib.rc_vector_to_color(c.h, vp(fl), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), vp(img), C.c_size_t(w * 3), None, C.c_int(0)))
    c._chk(c.lib.rc_shear_rate_to_color(c.h, vp(fl), C.c_size_t(w * 8), C.c_int(w), C.c_int(h), vp(img), C.c_size_t(w * 3), None, C.c_int(0)))
c.synchronize(); c.close()
