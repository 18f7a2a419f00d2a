import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
import tests.gpu_cases as c
from mmseg_b200 import kernels as K
from mmseg_b200.kernels import Blocked
DEV = c.DEV
torch.manual_seed(0)
heads, hd, n_tok, n_img = 4, 32, 300, 2
C = heads * hd
shape = (1, 1, n_tok)
def blocked(t):
    b = Blocked(n_img, C, *shape, False, DEV); K.pack_ncdhw(t.reshape(n_img, C, *shape).contiguous(), b); return b
q, k, v, do = (c._bf(torch.randn(n_img, C, n_tok, device=DEV)) for _ in range(4))
kvb = Blocked(n_img, 2 * C, *shape, False, DEV)
K.pack_ncdhw(torch.cat([k, v], 1).reshape(n_img, 2 * C, *shape).contiguous(), kvb)
qb, dob = blocked(q), blocked(do)
ob = Blocked(n_img, C, *shape, False, DEV)
lse = torch.empty((n_img, heads, n_tok), dtype=torch.float32, device=DEV)
scale = float(hd) ** -0.5
print("fwd", flush=True)
K.cross_attention(qb, 0, kvb, 0, C, ob, 0, heads, hd, scale, lse=lse)
torch.cuda.synchronize(); print("fwd ok", lse[0,0,:4], flush=True)
dqb = Blocked(n_img, C, *shape, False, DEV); dkvb = Blocked(n_img, 2 * C, *shape, False, DEV)
print("bwd", flush=True)
dbg = torch.zeros(24 * 32, dtype=torch.int32).pin_memory()
os.environ["MMSEG_ATTN_DBG_PTR"] = str(dbg.data_ptr())
K.cross_attention_bwd(qb, 0, kvb, 0, C, ob, 0, dob, 0, lse, dqb, 0, dkvb, 0, C, heads, hd, scale)
time.sleep(6)
for ct in range(24):
    d = dbg[ct * 32:(ct + 1) * 32]
    print(ct, "prod", d[0:2].tolist(), "mma", d[8:13].tolist(), "soft", d[16:24].tolist(), flush=True)
os._exit(0)
