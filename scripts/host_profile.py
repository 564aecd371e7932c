import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
from oracle import generators as G
from rec_now_b200.rec_block import pairwise_loss_from_batch as PW
d = G.cfg3(0)
s, y, w = (torch.tensor(d[k]).cuda() for k in ("s", "y", "w"))
g = torch.tensor(d["g"]).cuda()
def api():
    lg = s.detach().requires_grad_(True)
    loss = PW.pairwise_loss(lg, y, g, click_occurance_power=-0.5, label_pair_to_weight_func=PW.label_gain_times_sample_weight, sample_weight=w)
    loss.backward()
for _ in range(50): api()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(1000): api()
pr.disable(); torch.cuda.synchronize()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28); print(st.getvalue()[:6000])
