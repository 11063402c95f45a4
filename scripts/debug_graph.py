import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
dev = torch.device('cuda:0')
def mk(graph, seed=11):
    random.seed(seed)
    algo, kfs, cur = bench.build_algorithm(dev, seed=seed)
    algo.config.graph_mapping = graph
    algo.config.min_sample_pixels = 256
    return algo, kfs + [cur]
ag, fg = mk(True); ae, fe = mk(False)
n = 4
ag.setup_optimizers(n, fg, True); sess = ag.mapping_session(fg); sess.begin(fg)
opt = ae.setup_optimizers(n, fe, True)
tg, te = ag.model.embed_fn.params, ae.model.embed_fn.params
sgs = lambda: ag.model_optimizers.optimizers['embed_fn'].state[tg]
for i in range(n):
    mg0 = sgs()['exp_avg'].clone() if sgs() else torch.zeros_like(tg)
    vg0 = sgs()['exp_avg_sq'].clone() if sgs() else torch.zeros_like(tg)
    pg0 = tg.detach().clone()
    torch.manual_seed(100 + i); lg = float(sess.step(i, fg))
    torch.manual_seed(100 + i)
    opt.zero_grad_all(); le = ae.get_loss(fe, True, i, n); le.backward()
    gt_e = ae.model.embed_fn.params.grad; gt_g = sess.grads[ag.model.embed_fn.params]
    gd = [(p.grad - sess.grads[q]).abs().max().item() / (p.grad.abs().max().item() + 1e-30) for p, q in zip(ae.model.decoder.parameters(), ag.model.decoder.parameters())]
    opt.optimizer_step_all(step=i)
    gg = sess.grads[tg]
    m_exp = mg0 + (gg - mg0) * 0.1
    v_exp = vg0 * 0.99 + 0.01 * gg * gg
    t = i + 1
    p_exp = pg0 - (1e-2 / (1 - 0.9**t)) * m_exp / (v_exp.sqrt() / (1 - 0.99**t)**0.5 + 1e-15)
    print('   graph-path: m vs expected', (sgs()['exp_avg'] - m_exp).abs().max().item(), 'v', (sgs()['exp_avg_sq'] - v_exp).abs().max().item(), 'p', (tg - p_exp).abs().max().item(), 'dyn', sess.dyn.cpu()[32:56].view(torch.float32).tolist())
    pt = (ae.model.embed_fn.params - ag.model.embed_fn.params).abs()
    pd = [(p - q).abs().max().item() for p, q in zip(ae.model.decoder.parameters(), ag.model.decoder.parameters())]
    se, sg = opt.optimizers['embed_fn'].state[ae.model.embed_fn.params], ag.model_optimizers.optimizers['embed_fn'].state[ag.model.embed_fn.params]
    print(i, 'loss', lg, float(le), 'gtab rel', ((gt_e - gt_g).abs().max() / gt_e.abs().max()).item(), 'gdec', gd,
          'ptab max', pt.max().item(), 'frac>1e-3', (pt > 1e-3).float().mean().item(), 'pdec', pd,
          'm diff', (se['exp_avg'] - sg['exp_avg']).abs().max().item(), 'step', float(se['step']), float(sg['step']))
