"""FLAGS access for the host layer.

The reference reads one global `absl.flags.FLAGS` namespace everywhere — in
module constructors *and* in forwards (reference config/config.py:1-127,
network/fs_net_repo/FaceRecon.py:15-16,37,114, PoseNet9D.py:27).  When the
reference's flags are defined (drop-in use under engine/train.py or
evaluation/evaluate.py) we read that very object, at the same moments.  Stand
alone (tests, bench.py on the GPU box — the reference tree does not travel) a
namespace with the reference's default values is used instead.
"""
import types

# defaults restated from reference config/config.py (only what the hot path reads)
_DEFAULTS = dict(
    obj_c=6, feat_c_R=1286, R_c=4, feat_c_ts=1289, Ts_c=6, feat_face=768,
    face_recon_c=30, gcn_sup_num=7, gcn_n_num=20, random_points=1028, train=1,
    aug_pc_pro=0.2, aug_pc_r=0.2, aug_rt_pro=0.3, aug_bb_pro=0.3, aug_bc_pro=0.3,
    fsnet_loss_type="l1", rot_1_w=8.0, rot_2_w=8.0, rot_regular=4.0, tran_w=8.0, size_w=8.0,
    recon_w=8.0, r_con_w=1.0, lr=1e-4, lr_pose=1.0, sample_method="basic",
    recon_n_w=3.0, recon_d_w=3.0, recon_v_w=1.0, recon_s_w=0.3, recon_f_w=1.0, recon_bb_r_w=1.0,
    recon_bb_t_w=1.0, recon_bb_s_w=1.0, recon_bb_self_w=1.0, recon_c_w=0.0, geo_p_w=1.0, geo_f_w=0.1,
    prop_pm_w=2.0, prop_sym_w=1.0, prop_r_reg_w=1.0,
)

_standalone = types.SimpleNamespace(**_DEFAULTS)


def get_flags():
    """The reference's absl FLAGS when they are defined and parsed, else defaults."""
    try:
        import absl.flags as flags
        F = flags.FLAGS
        if "gcn_n_num" in F and F.is_parsed():
            return F
    except Exception:
        pass
    return _standalone


class _Proxy:
    """`FLAGS.x` resolves at access time (the reference mutates FLAGS.train
    between construction and forward, evaluation/evaluate.py:39)."""

    def __getattr__(self, name):
        return getattr(get_flags(), name)

    def __setattr__(self, name, value):
        setattr(get_flags(), name, value)


FLAGS = _Proxy()
