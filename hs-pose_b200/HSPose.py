"""Drop-in replacement for the reference's network/HSPose.py (HSPose.py:23-275).

Same constructor `HSPose(train_stage)`, same `forward(**kwargs)` keyword set,
the same 16-key `output_dict` and 4-group `loss_dict`, `build_params`, and the
same `state_dict` keys (posenet.*), so engine/train.py:43-110 and
evaluation/evaluate.py:58-104 run unchanged and published checkpoints load
with strict=True.  The feature extractor underneath is the sm_100a kernel
stack (gcn3d / FaceRecon / PoseNet9D of this package).

Losses: all four groups (fs_net, recon_6face, geo, prop — the 19 scalar terms
engine/train.py:96-97 sums) are native restatements in losses.py; nothing is
imported from the reference tree.  An optional Chamfer(recon, PC) term
(BASELINE.json config 3) is added under fsnet_loss['Chamfer'] when `chamfer_w > 0`.
"""
import torch
import torch.nn as nn

from . import augment, ops
from .PoseNet9D import PoseNet9D
from .pc_sample import PC_sample
from .flags import FLAGS
from .losses import (chamfer_recon_loss, fs_net_loss, geo_transform_loss, get_gt_v, prop_rot_loss,
                     recon_6face_loss)


def control_loss(train_stage):
    """Which terms are active per stage (reference engine/organize_loss.py:1-14)."""
    if train_stage == 'PoseNet_only':
        return (['Rot1', 'Rot2', 'Rot1_cos', 'Rot2_cos', 'Rot_regular', 'Tran', 'Size', 'R_con'],
                ['Per_point', 'Point_voting'], ['Geo_point'], ['Prop_pm', 'Prop_sym'])
    if train_stage == 'FSNet_only':
        return ['Rot1', 'Rot2', 'Tran', 'Size', 'Recon'], [], [], []
    raise NotImplementedError


class HSPose(nn.Module):
    def __init__(self, train_stage, chamfer_w=0.0, loss_groups=("fsnet", "recon", "geo", "prop")):
        """`loss_groups` (extension, default = the reference's behaviour): which of the four loss groups
        are evaluated; a group left out comes back as an empty dict."""
        super(HSPose, self).__init__()
        self.posenet = PoseNet9D()
        self.train_stage = train_stage
        self.chamfer_w = chamfer_w
        self.loss_fs_net = fs_net_loss()
        self.loss_recon = recon_6face_loss()
        self.loss_geo = geo_transform_loss()
        self.loss_prop = prop_rot_loss()
        self.name_fs_list, self.name_recon_list, \
            self.name_geo_list, self.name_prop_list = control_loss(self.train_stage)
        self.loss_groups = tuple(loss_groups)

    def forward(self, PC=None, depth=None, obj_id=None, camK=None,
                gt_R=None, gt_t=None, gt_s=None, mean_shape=None, gt_2D=None, sym=None, aug_bb=None,
                aug_rt_t=None, aug_rt_r=None, def_mask=None, model_point=None, nocs_scale=None,
                do_loss=False):
        output_dict = {}
        if PC is None:
            # reference HSPose.py:40-50: the cloud is sampled from the depth ROI (K11, pc_sample.PC_sample)
            if self.train_stage != 'PoseNet_only':
                raise NotImplementedError
            if depth is None or def_mask is None or camK is None or gt_2D is None:
                raise ValueError("HSPose.forward needs PC, or depth + def_mask + camK + gt_2D to sample it from")
            FLAGS.sample_method = 'basic'
            # the reference draws (and never uses) a `sketch` tensor here; the draw is kept so that the CUDA
            # generator is in the same state when data_augment consumes it
            torch.rand([depth.shape[0], 6, depth.shape[2], depth.shape[3]], device=depth.device)
            PC = PC_sample(def_mask, depth, camK, gt_2D)
            if PC is None:
                return output_dict, None

        PC = PC.detach()
        if FLAGS.train:
            with torch.no_grad():
                PC, gt_R, gt_t, gt_s = self.data_augment(PC, gt_R, gt_t, gt_s, mean_shape, sym, aug_bb,
                                                         aug_rt_t, aug_rt_r, model_point, nocs_scale, obj_id)

        recon, face_normal, face_dis, face_f, p_green_R, p_red_R, f_green_R, f_red_R, \
            Pred_T, Pred_s = self.posenet(PC, obj_id)

        output_dict.update(mask=None, sketch=None, recon=recon, PC=PC, face_normal=face_normal,
                           face_dis=face_dis, face_f=face_f, p_green_R=p_green_R, p_red_R=p_red_R,
                           f_green_R=f_green_R, f_red_R=f_red_R, Pred_T=Pred_T, Pred_s=Pred_s,
                           gt_R=gt_R, gt_t=gt_t, gt_s=gt_s)
        if not do_loss:
            return output_dict

        if self.train_stage == 'Backbone_only':
            gt_green_v, gt_red_v = None, None
        else:
            gt_green_v, gt_red_v = get_gt_v(gt_R)
        pred_fsnet_list = {'Rot1': p_green_R, 'Rot1_f': f_green_R, 'Rot2': p_red_R, 'Rot2_f': f_red_R,
                           'Recon': recon, 'Tran': Pred_T, 'Size': Pred_s}
        gt_fsnet_list = {'Rot1': gt_green_v, 'Rot2': gt_red_v, 'Recon': PC, 'Tran': gt_t, 'Size': gt_s}
        if self._fused_losses_ok(PC, recon):
            return output_dict, self._fused_losses(PC, recon, p_green_R, p_red_R, f_green_R, f_red_R, Pred_T, Pred_s,
                                                   gt_R, gt_t, gt_s, mean_shape, sym, obj_id)

        fsnet_loss = {}
        if "fsnet" in self.loss_groups:
            fsnet_loss = self.loss_fs_net(self.name_fs_list, pred_fsnet_list, gt_fsnet_list, sym)
        if self.chamfer_w > 0 and recon is not None:
            fsnet_loss['Chamfer'] = chamfer_recon_loss(recon, PC, self.chamfer_w)

        prop_loss, recon_loss, geo_loss = {}, {}, {}
        if "prop" in self.loss_groups:
            pred_prop_list = {'Recon': recon, 'Rot1': p_green_R, 'Rot2': p_red_R, 'Tran': Pred_T,
                              'Scale': Pred_s, 'Rot1_f': f_green_R.detach(), 'Rot2_f': f_red_R.detach()}
            gt_prop_list = {'Points': PC, 'R': gt_R, 'T': gt_t, 'Mean_shape': mean_shape}
            prop_loss = self.loss_prop(self.name_prop_list, pred_prop_list, gt_prop_list, sym)
        if "recon" in self.loss_groups:
            pred_recon_list = {'F_n': face_normal, 'F_d': face_dis, 'F_c': face_f, 'Rot1': p_green_R,
                               'Rot1_f': f_green_R.detach(), 'Rot2': p_red_R, 'Rot2_f': f_red_R.detach(),
                               'Tran': Pred_T, 'Size': Pred_s}
            gt_recon_list = {'R': gt_R, 'T': gt_t, 'Size': gt_s, 'Mean_shape': mean_shape, 'Points': PC}
            recon_loss = self.loss_recon(self.name_recon_list, pred_recon_list, gt_recon_list, sym, obj_id)
        if "geo" in self.loss_groups:
            pred_geo_list = {'Rot1': p_green_R, 'Rot2': p_red_R, 'Tran': Pred_T, 'Size': Pred_s,
                             'Rot1_f': f_green_R.detach(), 'Rot2_f': f_red_R.detach()}
            gt_geo_list = {'Points': PC, 'R': gt_R, 'T': gt_t, 'Mean_shape': mean_shape}
            geo_loss = self.loss_geo(self.name_geo_list, pred_geo_list, gt_geo_list, sym)

        loss_dict = {'fsnet_loss': fsnet_loss, 'recon_loss': recon_loss, 'geo_loss': geo_loss,
                     'prop_loss': prop_loss}
        return output_dict, loss_dict

    # ---- K8: the whole 19-term loss graph as two launches each way (ops.fused_losses)
    fused_losses = True    # set False to evaluate the losses with the tensor-algebra modules of losses.py
    _loss_mask = {}        # (selection, device) -> 0/1 vector over ops.LOSS_TERMS

    def _fused_losses_ok(self, PC, recon):
        return (self.fused_losses and PC.is_cuda and recon is not None and self.train_stage == 'PoseNet_only'
                and FLAGS.fsnet_loss_type == 'l1' and FLAGS.prop_sym_w > 0
                and getattr(self.posenet, "face_raw", None) is not None)

    def _fused_losses(self, PC, recon, p_green_R, p_red_R, f_green_R, f_red_R, Pred_T, Pred_s, gt_R, gt_t, gt_s,
                      mean_shape, sym, obj_id):
        weights = [getattr(FLAGS, n) for n in ops.LOSS_WEIGHT_FLAGS]
        t = ops.fused_losses(weights, self.posenet.face_raw, recon, p_green_R, p_red_R, f_green_R, f_red_R, Pred_T,
                             Pred_s, PC, gt_R, gt_t, gt_s, mean_shape, sym, obj_id)
        g = self.loss_groups
        fs = {k: t[k] for k in ("Rot1", "Rot1_cos", "Rot2", "Rot2_cos", "Rot_r_a", "Tran", "Size", "R_con")} \
            if "fsnet" in g else {}
        if self.chamfer_w > 0:
            fs['Chamfer'] = chamfer_recon_loss(recon, PC, self.chamfer_w)
        rec = {k: t[k] for k in ("recon_per_p", "recon_p_f", "recon_point_vote", "recon_point_r", "recon_point_t",
                                 "recon_point_s", "recon_point_self")} if "recon" in g else {}
        geo = {"geo_point": t["geo_point"]} if "geo" in g else {}
        prop = {k: t[k] for k in ("Prop_pm", "Prop_sym_recon", "Prop_sym_rt")} if "prop" in g else {}
        out = ops.LossGroups(fsnet_loss=fs, recon_loss=rec, geo_loss=geo, prop_loss=prop)
        # the sum of the selected terms as ONE masked reduction of the term vector (+ Chamfer)
        sel = [1.0 if any(name in grp for grp in (fs, rec, geo, prop)) else 0.0 for name in ops.LOSS_TERMS]
        mask = self._loss_mask.get((tuple(sel), t.vector.device))
        if mask is None:
            mask = self._loss_mask[(tuple(sel), t.vector.device)] = torch.tensor(sel, device=t.vector.device)
        out.total = (t.vector * mask).sum()
        if 'Chamfer' in fs:
            out.total = out.total + fs['Chamfer'].reshape(()).float()
        return out

    def data_augment(self, PC, gt_R, gt_t, gt_s, mean_shape, sym, aug_bb, aug_rt_t, aug_rt_r,
                     model_point, nocs_scale, obj_ids, check_points=False):
        """Reference HSPose.py:185-256: four Bernoulli-gated deformations; the device RNG is
        consumed in the same order (prob_bb, prob_rt, prob_bc, ey_up, ey_down, prob_pc, defor)."""
        bs = PC.shape[0]
        if PC.is_cuda and not check_points:
            # K10: one launch; the uniform draws keep the reference's order and shapes
            dev = PC.device
            g_bb, g_rt, g_bc = (torch.rand((bs, 1), device=dev) for _ in range(3))
            ey_up, ey_down = torch.rand((bs, 1), device=dev), torch.rand((bs, 1), device=dev)
            g_pc = torch.rand((bs, 1), device=dev)
            defor = torch.rand(PC.shape, device=dev)
            return ops.augment(PC, gt_R, gt_t, gt_s, mean_shape, sym, aug_bb, aug_rt_t, aug_rt_r, model_point,
                               nocs_scale, obj_ids, torch.cat([g_bb, g_rt, g_bc, g_pc], dim=1),
                               torch.cat([ey_up, ey_down], dim=1), defor,
                               (FLAGS.aug_bb_pro, FLAGS.aug_rt_pro, FLAGS.aug_bc_pro, FLAGS.aug_pc_pro), FLAGS.aug_pc_r)
        flag = torch.rand((bs, 1), device=PC.device) < FLAGS.aug_bb_pro
        PC_new, s_new, mp_new = augment.deform_bb(PC, model_point, gt_R, gt_t, gt_s + mean_shape, sym, aug_bb)
        PC = torch.where(flag.unsqueeze(-1), PC_new, PC)
        gt_s = torch.where(flag, s_new - mean_shape, gt_s)
        model_point = torch.where(flag.unsqueeze(-1), mp_new, model_point)

        flag = torch.rand((bs, 1), device=PC.device) < FLAGS.aug_rt_pro
        PC_new, R_new, t_new = augment.deform_rt(PC, gt_R, gt_t, aug_rt_t, aug_rt_r)
        PC = torch.where(flag.unsqueeze(-1), PC_new, PC)
        gt_R = torch.where(flag.unsqueeze(-1), R_new, gt_R)
        gt_t = torch.where(flag, t_new, gt_t)

        flag = torch.logical_and(torch.rand((bs, 1), device=PC.device) < FLAGS.aug_bc_pro,
                                 torch.logical_or(obj_ids == 5, obj_ids == 1).unsqueeze(-1))
        PC_new, s_new, _, _ = augment.deform_bc(PC, gt_R, gt_t, gt_s + mean_shape, model_point, nocs_scale)
        PC = torch.where(flag.unsqueeze(-1), PC_new, PC)
        gt_s = torch.where(flag, s_new - mean_shape, gt_s)

        flag = torch.rand((bs, 1), device=PC.device) < FLAGS.aug_pc_pro
        PC_new, _ = augment.deform_pc(PC, gt_t, FLAGS.aug_pc_r)
        PC = torch.where(flag.unsqueeze(-1), PC_new, PC)
        return PC, gt_R, gt_t, gt_s

    def build_params(self, training_stage_freeze=None):
        """Reference HSPose.py:258-275."""
        if training_stage_freeze and 'pose' in training_stage_freeze:
            for param in self.posenet.parameters():
                param.requires_grad = False
        return [{"params": filter(lambda p: p.requires_grad, self.posenet.parameters()),
                 "lr": float(FLAGS.lr) * FLAGS.lr_pose}]
