import ctypes as C, os, sys
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import baby_shark_b200 as bs
from baby_shark_b200 import synth
tris, vs, desc = synth.config_mesh(5, 1.0)
L, ctx = bs.load_library(), bs.Context.default()
d = torch.from_numpy(tris).cuda()
for world, rank in ((4, 1), (8, 3), (8, 0)):
    for rep in range(2):
        h = C.c_void_p()
        ctx.check(L.bs_mesh_to_volume_sharded(ctx._h, C.c_void_p(d.data_ptr()), tris.shape[0], vs, 0, rank, world, C.byref(h)))
        st = ctx.last_stats(); L.bs_volume_free(h)
    print(world, rank, {k: round(v, 2) for k, v in st.items() if 'sign' in k or 'heavy' in k or k in ('n_bricks', 'n_bricks_owned')})
