"""LossCalculator: the drop-in replacement for the reference's `model/_loss.py::LossCalculator`.

Same constructor keywords, same `forward(stu_out, tea_out, model_type) -> (loss, dict)`, same
`get_control_output / set_percent / set_scale`, same `.loss` ModuleDict, same weighting rules
(reference model/_loss.py:18-55, 100-116, 118-216), so `DistillModel` (reference model/distil_model.py:51-52,
100,112) and `DualDistillModel` (model/dual_distill_model.py:79-80,124,131) can call it unchanged.
Underneath, each tower's losses run as one-pass CUDA kernels with a single deterministic finalize, and the
two-tower hard/soft-label terms run as the fused tcgen05 contrastive kernel straight from the embeddings
(`last_representation`), never touching the materialised B x B logits.

In scope: hard_label, soft_label, attention_probs_kl, hidden_rep_mse, embedding_mse; widened
(section 8f) to the losses of the three shipped configs and their siblings: out_l1, out_cos, cos_diff,
attention_probs_mse, attention_score_mse, out_kl, out_ce, logits_mse, last_value_map_kl.  The reference's remaining names
(vit_kd, fine_grain, smd) are recognised but raise NotImplementedError.
"""
from typing import Dict, List, Union

import torch
from torch import nn

from .. import contrastive, ops
from .._lib import DistillClipB200Error
from .component.output import CLIPOutput, ControlOutput, TextTransformerOutput, VisionTransformerOutput
from .loss_component import (AttentionProbsKL, AttentionProbsMSE, AttentionScoreMSE, CLIPCosDiff, EmbedMSELoss,
                             HardLabel, HiddenMSE, LastValueMapKL, LogitsMSE, OutCELoss, OutCosLoss, OutKLLoss, OutL1Loss, SoftLabel)

# reference _loss.py:9-12 -- including the missing comma that fuses 'smd' and 'hard_label' 
LOSSNAME = ['out_l1', 'out_ce', 'out_kl', 'out_cos', 'embedding_mse', 'attention_score_mse',
            'attention_probs_mse', 'hidden_rep_mse', 'attention_probs_kl', 'last_value_map_kl',
            'vit_kd', 'smd' 'hard_label', 'soft_label', 'fine_grain', 'logits_mse']
IMAGE_TEXT_LOSS = ['hard_label', 'soft_label', 'logits_mse', 'fine_grain', 'cos_diff']

# names the reference accepts (_loss.py:60-94) that are outside this build's hot-path scope
_REFERENCE_ONLY = ('vit_kd', 'fine_grain', 'smd')

# one-tower losses: name -> (kernel family, student field, is a list of layers)
_TOWER_KERNELS = {
    'out_l1': (ops.KIND_L1, 'last_representation', False),
    'out_cos': (ops.KIND_COS, 'last_representation', False),
    'embedding_mse': (ops.KIND_MSE, 'embedding', False),
    'attention_score_mse': (ops.KIND_ATTN_MSE, 'attention_scores', True),
    'attention_probs_mse': (ops.KIND_ATTN_MSE, 'attention_probs', True),
    'hidden_rep_mse': (ops.KIND_MSE, 'representations', True),
    'attention_probs_kl': (ops.KIND_ATTN_KL, 'attention_probs', True),
}


class LossCalculator(nn.Module):
    #: use the fused embeddings->loss kernel for hard/soft label in two-tower mode (else the logits modules)
    fused_contrastive = True
    #: upstream gradient assumed by the one-pass streaming kernels; None = distillclip_b200.ops.EXPECTED_GRAD_SCALE (1.0).
    #: A different upstream value is always honoured exactly (backward recomputes the gradients with the true value, one
    #: more pass); under fp16 AMP set `grad_scaler` instead so that the extra pass never happens.
    expected_grad_scale = None
    #: a torch.amp.GradScaler: its device-side scale tensor is read by the forward kernels, so the one-pass gradients are
    #: written -- and rounded to fp16 exactly once -- at the magnitude `scaler.scale(loss).backward()` will ask for
    grad_scaler = None
    #: torch.distributed process group for GLOBAL-batch contrastive losses (embeddings all-gathered, each rank
    #: computes its row slice; SURVEY.md F5: the reference itself is local-batch, so the default is None)
    contrastive_group = None
    _warned_fp16 = False

    def __init__(self, loss_name: List, loss_scale: dict = None,
                 temperature=None, percent=None, smd_tau: float = 0.04, vit_kd_para: Dict = None):
        super().__init__()
        self.loss_name = loss_name
        given_scale = {n: 1 for n in loss_name} if loss_scale is None else loss_scale
        self.loss_scale = {n: given_scale.get(n, 1) for n in loss_name}

        # percent rules of reference _loss.py:29-42 (the fill value divides by the number of GIVEN entries)
        if percent is None:
            percent = {n: 1 / len(loss_name) for n in loss_name}
        self.percent = percent
        fill = (1 - sum(percent.values())) / len(percent)
        if len(loss_name) != len(percent.keys()) and fill <= 0:
            raise ValueError(
                f"there are some loss default percent is negative. Please check the sum of the percent {percent}"
                f"the default_value is {fill} = (1 - sum(percent.values())) / len(percent)")
        for n in loss_name:
            percent.setdefault(n, fill)
        assert abs(sum(percent.values()) - 1) <= 1e-5

        self.temperature = temperature
        if vit_kd_para is not None:
            vit_kd_para.setdefault('low_layers_num', 2)
            vit_kd_para.setdefault('high_layers_num', 1)
        self.vit_kd_para = vit_kd_para
        self.smd_tau = smd_tau
        self.loss = self._init_loss()

        print(self.percent)
        print(self.loss_scale)

    def _init_loss(self):
        table = {
            'out_l1': OutL1Loss, 'out_cos': OutCosLoss, 'embedding_mse': EmbedMSELoss,
            'attention_score_mse': AttentionScoreMSE, 'attention_probs_mse': AttentionProbsMSE,
            'hidden_rep_mse': HiddenMSE, 'attention_probs_kl': AttentionProbsKL,
            'hard_label': HardLabel, 'soft_label': lambda: SoftLabel(self.temperature), 'cos_diff': CLIPCosDiff,
            'out_ce': OutCELoss, 'out_kl': lambda: OutKLLoss(self.temperature), 'logits_mse': LogitsMSE,
            'last_value_map_kl': LastValueMapKL,
        }
        losses = nn.ModuleDict()
        for n in self.loss_name:
            if n in table:
                losses[n] = table[n]()
            elif n in _REFERENCE_ONLY:
                raise NotImplementedError(
                    f"loss '{n}' exists in the reference but is outside the B200 hot-path build "
                    f"(in scope: {sorted(table)})")
            else:
                raise ValueError("Invalid Loss Type!")
        return losses

    def get_control_output(self):
        need_para = ControlOutput()
        flags = {'embedding_mse': 'need_emb', 'attention_score_mse': 'need_attn_score',
                 'attention_probs_mse': 'need_attn_prob', 'hidden_rep_mse': 'need_rep',
                 # reference _loss.py:111-112 sets this non-existent attribute instead of need_attn_prob
                 # (SURVEY.md F8); preserved so encoders see exactly what they saw before
                 'attention_probs_kl': 'attention_probs_mse',
                 'last_value_map_kl': 'need_value_map'}
        for n in self.loss_name:
            if n in flags:
                setattr(need_para, flags[n], True)
        return need_para

    # ------------------------------------------------------------------------------------------
    def _assumed_upstream(self):
        """-> (host float, optional device scalar): the upstream gradient of the TOTAL loss the one-pass kernels assume."""
        expected = float(self.expected_grad_scale if self.expected_grad_scale is not None else ops.EXPECTED_GRAD_SCALE)
        return expected, ops.grad_scaler_mult(self.grad_scaler)

    def cal_one_tower_loss(self,
                           stu_out: Union[VisionTransformerOutput, TextTransformerOutput],
                           tea_out: Union[VisionTransformerOutput, TextTransformerOutput], tower_weight: float = 1.0):
        """reference _loss.py:155-202: raw values per name, `* scale`, `loss += value * percent`.  `tower_weight`: the factor
        the caller applies to this tower's total (0.5 in two-tower mode), folded into the assumed upstream gradient."""
        expected, fwd_mult = self._assumed_upstream()
        expected *= float(tower_weight)
        raw_python = {}          # names whose value is a python number (empty teacher list edge case)
        module_res = {}          # row-softmax losses on the pooled outputs (own small kernels)
        for name in self.loss:
            if name == 'out_kl':
                assert self.temperature, 'You should give the temperature for the kl loss'
            if name in ('out_ce', 'out_kl'):
                module_res[name] = self.loss[name](stu_out.last_representation, tea_out.last_representation)
            elif name == 'last_value_map_kl':
                w = float(self.loss_scale.get(name, 1)) * float(self.percent.get(name, 0))
                module_res[name] = ops.value_map_kl(stu_out.value_map, tea_out.value_map, expected=w * expected if w else None,
                                                    fwd_mult=fwd_mult if w else None)
        spec, tensors, order = [], [], []
        for name in self.loss:
            if name not in _TOWER_KERNELS:
                continue
            kind, field, is_list = _TOWER_KERNELS[name]
            stu, tea = getattr(stu_out, field), getattr(tea_out, field)
            if not is_list:
                stu, tea = [stu], [tea]
            divisor = len(stu)
            if divisor == 0:
                raise ZeroDivisionError("division by zero")
            s, t = ops._prep_pair(stu, tea)
            if not s:
                raw_python[name] = 0.0
                continue
            order.append(name)
            spec.append((kind, divisor, len(s), name))
            tensors.append((s, t))
        if fwd_mult is None and expected == float(tower_weight) and not self._warned_fp16 and any(
                x.dtype == torch.float16 and x.requires_grad for s, _ in tensors for x in s):
            self._warned_fp16 = True
            import warnings
            warnings.warn("distillclip_b200: fp16 student tensors with an assumed upstream gradient of 1.0 -- under a GradScaler "
                          "the gradients are recomputed in backward at the true loss scale (exact, one extra pass over the "
                          "inputs); set LossCalculator.grad_scaler = scaler to write them once")

        weights_ok = all(n in self.loss_scale and n in self.percent for n in order)
        cal_res = {}
        fused_total, fused = None, {}
        if order and weights_ok:
            full_spec = [(kind, div, n, float(self.loss_scale[name]), float(self.percent[name]))
                         for kind, div, n, name in spec]
            flat = [x for s, t in tensors for x in (*s, *t)]
            outs = ops.TowerLossFn.apply(full_spec, expected, fwd_mult, *flat)
            fused_total, fused = outs[0], dict(zip(order, outs[1:]))
        else:
            # after set_scale / set_percent with a partial dict: raw values from one single-term launch each; the loop below
            # scales and adds only what the reference's loop over loss_scale would
            for (kind, div, n, name), (s, t) in zip(spec, tensors):
                cal_res[name] = ops.TowerLossFn.apply([(kind, div, n, 1.0, 1.0)], float(ops.EXPECTED_GRAD_SCALE), None, *s, *t)[0]
        for name in self.loss:                      # dict order of the reference: order of self.loss
            if name in fused:
                cal_res[name] = fused[name]
            elif name in raw_python:
                cal_res[name] = raw_python[name]
            elif name in module_res:
                cal_res[name] = module_res[name]

        loss = 0
        pending_fused = fused_total is not None
        for (loss_name, scale) in self.loss_scale.items():
            if loss_name in IMAGE_TEXT_LOSS:
                continue
            if loss_name in fused:
                if pending_fused:                   # all fused names were weighted inside the kernel
                    loss = fused_total if (isinstance(loss, int) and loss == 0) else loss + fused_total
                    pending_fused = False
                continue
            cal_res[loss_name] = cal_res[loss_name] * scale
            loss = loss + cal_res[loss_name] * self.percent[loss_name]      # not in place: `loss` may be an output view of the tower kernel
        return loss, cal_res

    def cal_tow_tower_loss(self, stu_out: CLIPOutput, tea_out: CLIPOutput):
        """reference _loss.py:118-153."""
        cal_res = {}
        # each tower's total enters the loss as 0.5 * total (reference _loss.py:148): tell the one-pass kernels
        image_loss, image_loss_dict = self.cal_one_tower_loss(stu_out.visual_output, tea_out.visual_output, tower_weight=0.5)
        text_loss, text_loss_dict = self.cal_one_tower_loss(stu_out.text_output, tea_out.text_output, tower_weight=0.5)
        for k, v in image_loss_dict.items():
            cal_res['image_' + k] = v
        for k, v in text_loss_dict.items():
            cal_res['text_' + k] = v

        want_hard, want_soft = 'hard_label' in self.loss, 'soft_label' in self.loss
        want_cos, want_mse = 'cos_diff' in self.loss, 'logits_mse' in self.loss
        if want_soft:
            assert self.temperature
        stu_img, stu_txt = stu_out.visual_output.last_representation, stu_out.text_output.last_representation
        need_tea = want_soft or want_cos or want_mse
        tea_img = tea_out.visual_output.last_representation if need_tea else None
        tea_txt = tea_out.text_output.last_representation if need_tea else None
        fused = {}
        if (want_hard or want_soft or want_cos or want_mse) and self.fused_contrastive and contrastive.fused_supported(
                stu_img, stu_txt, self.temperature if want_soft else None, tea_img, tea_txt):
            # cos_diff / logits_mse ride on the pipeline kernels only; otherwise they stay on the caller's logits
            extras = (want_cos or want_mse) and contrastive.extras_supported(stu_img, self.contrastive_group)
            f_cos, f_mse = want_cos and extras, want_mse and extras
            # scale / percent of reference _loss.py:231-234 applied on the device; a name missing from loss_scale (after
            # set_scale with a partial dict) is neither scaled nor added, exactly like the reference's loop over loss_scale
            def weight(name, on):
                if not on or name not in self.loss_scale:
                    return 1.0, 0.0
                return float(self.loss_scale[name]), float(self.percent[name])
            ws = [weight('hard_label', want_hard), weight('soft_label', want_soft), weight('cos_diff', f_cos), weight('logits_mse', f_mse)]
            if want_hard or want_soft or f_cos or f_mse:
                fused = contrastive.clip_contrastive(stu_img, stu_txt, tea_img if (want_soft or f_cos or f_mse) else None,
                                                     tea_txt if (want_soft or f_cos or f_mse) else None,
                                                     self.temperature if want_soft else None, want_hard, want_soft,
                                                     group=self.contrastive_group, percent=[w[1] for w in ws],
                                                     scale=[w[0] for w in ws], want_cos_diff=f_cos, want_logits_mse=f_mse)

        def logits_of(out, who):
            if out.i2t_logits is None or out.t2i_logits is None:
                raise DistillClipB200Error(
                    f"{who} CLIPOutput carries no logits (LazyLogitsCLIP?) but '{loss_name}' cannot run on the fused path here: "
                    "it needs CUDA bf16/fp16 [B, D] embeddings with D % 8 == 0, a teacher of the same shape and "
                    f"temperature >= {contrastive.MIN_FUSED_TEMPERATURE}; materialise the logits or fix the inputs")
            return out.i2t_logits, out.t2i_logits
        for loss_name in self.loss_name:
            if loss_name in ('cos_diff', 'logits_mse') and loss_name in fused:
                cal_res[loss_name] = fused[loss_name]
                continue
            if loss_name in ('cos_diff', 'logits_mse'):     # reference _loss.py:138-145, on the caller's materialised logits
                loss = self.loss[loss_name]
                (s_i2t, s_t2i), (t_i2t, t_t2i) = logits_of(stu_out, 'student'), logits_of(tea_out, 'teacher')
                cal_res[loss_name] = 0.5 * (loss(s_i2t, t_i2t) + loss(s_t2i, t_t2i))
                continue
            if loss_name not in ('hard_label', 'soft_label'):
                continue
            if loss_name in fused:
                cal_res[loss_name] = fused[loss_name]
            elif loss_name == 'hard_label':
                loss = self.loss[loss_name]
                s_i2t, s_t2i = logits_of(stu_out, 'student')
                cal_res[loss_name] = 0.5 * (loss(s_i2t) + loss(s_t2i))
            else:
                loss = self.loss[loss_name]
                (s_i2t, s_t2i), (t_i2t, t_t2i) = logits_of(stu_out, 'student'), logits_of(tea_out, 'teacher')
                cal_res[loss_name] = 0.5 * (loss(s_i2t, t_i2t) + loss(s_t2i, t_t2i))

        loss = 0.5 * (image_loss + text_loss)
        fused_pending = 'total' in fused
        for (loss_name, scale) in self.loss_scale.items():
            if loss_name not in IMAGE_TEXT_LOSS:
                continue
            if loss_name in fused:                          # already scaled; their weighted sum was formed on the device
                if fused_pending:
                    loss = loss + fused['total']
                    fused_pending = False
                continue
            cal_res[loss_name] = cal_res[loss_name] * scale
            loss += cal_res[loss_name] * self.percent[loss_name]
        return loss, cal_res

    def forward(self, stu_out: Union[CLIPOutput, VisionTransformerOutput, TextTransformerOutput],
                tea_out: Union[CLIPOutput, VisionTransformerOutput, TextTransformerOutput],
                model_type: str):
        if model_type == 'all':
            return self.cal_tow_tower_loss(stu_out, tea_out)
        return self.cal_one_tower_loss(stu_out, tea_out)

    def set_percent(self, new_percent):
        self.percent = new_percent

    def set_scale(self, new_scale):
        self.loss_scale = new_scale
