"""Debug aid: gradients of one train step under (mode, streams) configurations, saved to /tmp for comparison."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def run(tag, graph):
    import irfd_oracle as O
    import speak_hack_b200 as P
    from speak_hack_b200.trainer import IRFDTrainer

    dev = torch.device("cuda:0")
    torch.manual_seed(O.WEIGHT_SEED)
    net = P.IRFD().to(dev).train()
    net.Gd.style_mixing_prob = 0.0
    tr = IRFDTrainer(net, use_cuda_graph=graph)
    x_s, x_t = O.synthetic_pair(2)
    torch.manual_seed(321)
    torch.cuda.manual_seed(654)
    loss = tr.train_step(x_s.to(dev), x_t.to(dev))
    torch.cuda.synchronize()
    g = {"loss": loss.detach().cpu()}
    for en in ("Gd", "Ei", "Ee", "Ep"):
        for n, p in getattr(net, en).named_parameters():
            if p.grad is not None:
                g[f"{en}.{n}"] = p.grad.detach().cpu().clone()
    torch.save(g, f"/tmp/g_{tag}.pt")


def cmp(a, b):
    ga, gb = torch.load(f"/tmp/g_{a}.pt"), torch.load(f"/tmp/g_{b}.pt")
    worst, bad = ("", 0.0), 0
    for k, v in ga.items():
        if "noise" in k or k == "loss":
            continue
        e = float((gb[k].double() - v.double()).norm() / v.double().norm().clamp_min(1e-30))
        if e > 1e-4:
            bad += 1
        if e > worst[1]:
            worst = (k, e)
    print(f"{a} vs {b}: loss {float(ga['loss']):.6e} / {float(gb['loss']):.6e}; {bad} tensors > 1e-4; worst {worst}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], sys.argv[2] == "1")
    else:
        import subprocess

        for tag, graph, env in (("eager_ns", "0", "1"), ("eager_ns2", "0", "1"), ("eager_st", "0", ""), ("eager_st2", "0", ""),
                                ("graph_st", "1", ""), ("graph_st2", "1", "")):
            e = dict(os.environ)
            if env:
                e["IRFD_NO_ENC_STREAMS"] = "1"
            subprocess.run([sys.executable, __file__, tag, graph], env=e, check=True)
        for a, b in (("eager_ns", "eager_ns2"), ("eager_ns", "eager_st"), ("eager_st", "eager_st2"), ("eager_st", "graph_st"),
                     ("eager_ns", "graph_st"), ("graph_st", "graph_st2")):
            cmp(a, b)
