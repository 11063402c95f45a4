"""Host-side profile of the Co-SLAM e2e mapping iteration (cProfile), GPU box only."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device('cuda:0')
algo, kfs, cur = bench.build_algorithm(dev, seed=1234)
frames = kfs + [cur]
algo.config.min_sample_pixels = bench.MAP_CUR
K = 50
optim = algo.setup_optimizers(K, frames, is_mapping=True)


def step(i):
    optim.zero_grad_all()
    loss = algo.get_loss(frames, True, i, K)
    loss.backward()
    optim.optimizer_step_all(step=i)
    return loss.item()


for i in range(10):
    step(i)
torch.cuda.synchronize()
# coarse phase timers
T = dict(zero=0, input=0, fwd=0, bwd=0, opt=0, item=0)
for i in range(K):
    t0 = time.perf_counter(); optim.zero_grad_all()
    t1 = time.perf_counter(); algo.model.freeze_map_grads = False; inp = algo.get_model_input(frames, True)
    t2 = time.perf_counter(); out = algo.model(inp); ld = algo.model.get_loss_dict(out, inp, True, i); loss = sum(ld.values())
    t3 = time.perf_counter(); loss.backward()
    t4 = time.perf_counter(); optim.optimizer_step_all(step=i)
    t5 = time.perf_counter(); loss.item()
    t6 = time.perf_counter()
    for k, a, b in zip(T, (t0, t1, t2, t3, t4, t5), (t1, t2, t3, t4, t5, t6)):
        T[k] += (b - a) / K * 1e3
print('phase ms/iter:', {k: round(v, 3) for k, v in T.items()}, 'total', round(sum(T.values()), 3))
pr = cProfile.Profile()
pr.enable()
for i in range(K):
    step(i)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
