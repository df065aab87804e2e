import sys, os, numpy as np, torch
sys.path.insert(0,'.')
from hspose_b200.losses import recon_6face_loss
from hspose_b200.synth import synth_predictions
G=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'losses.npz'))
def run(dt):
    pred, gt = synth_predictions(12,257,seed=11)
    pred={k:v.to(dt) for k,v in pred.items()}; gt={k:v.to(dt) for k,v in gt.items()}
    p={k:v.clone().requires_grad_() for k,v in pred.items()}
    t=recon_6face_loss()(['Per_point','Point_voting'],
        {'F_n': p["face_normal"], 'F_d': p["face_dis"], 'F_c': p["face_f"], 'Rot1': p["p_green_R"],
             'Rot1_f': p["f_green_R"].detach(), 'Rot2': p["p_red_R"], 'Rot2_f': p["f_red_R"].detach(),
             'Tran': p["Pred_T"], 'Size': p["Pred_s"]},
            {'R': gt["gt_R"], 'T': gt["gt_t"], 'Size': gt["gt_s"], 'Mean_shape': gt["mean_shape"],
             'Points': gt["PC"]}, gt["sym"], gt["obj_id"])
    sum(t.values()).backward()
    return {k:v.grad for k,v in p.items() if v.grad is not None}, t
g32,t32=run(torch.float32); g64,t64=run(torch.float64)
for k in g64:
    key='recon::grad::'+k
    if key in G.files:
        ref=torch.from_numpy(G[key]).double()
        sc=g64[k].abs().max().item()
        print(k, 'scale %.3g'%sc, 'mine32-64 %.3g'%((g32[k].double()-g64[k]).abs().max().item()/sc), 'ref32-mine64 %.3g'%((ref-g64[k]).abs().max().item()/sc))
for k in t64:
    print(k, float(t64[k]), float(t32[k]), float(G['recon::'+k]))
