import torch, collections
ga, gb = torch.load("/tmp/g_eager_st.pt"), torch.load("/tmp/g_graph_st.pt")
grp = collections.defaultdict(list)
for k, v in ga.items():
    if k == "loss": continue
    e = float((gb[k].double() - v.double()).norm() / v.double().norm().clamp_min(1e-30))
    r = float(gb[k].double().norm() / v.double().norm().clamp_min(1e-30))
    grp[k.split(".")[0] + ("(noise)" if "noise" in k else "")].append((e, r, k))
for g, lst in grp.items():
    lst.sort()
    print(g, "n", len(lst), "median err %.3e max err %.3e (%s) norm-ratio of worst %.4f" % (lst[len(lst)//2][0], lst[-1][0], lst[-1][2], lst[-1][1]))
for k in ("Ei.7.2.bn3.bias", "Ee.7.2.bn3.bias", "Ep.7.2.bn3.bias", "Ei.7.2.conv3.weight", "Gd.mapping.0.weight", "Gd.mapping.7.weight"):
    v, w = ga[k].double(), gb[k].double()
    print(k, "err %.3e ratio %.4f cos %.6f" % (float((w - v).norm() / v.norm()), float(w.norm() / v.norm()), float((v * w).sum() / v.norm() / w.norm())))
