"""Drop-in ``ViBERTgridNet`` whose joint forward runs on hand-written sm_100a
kernels through the C-ABI library ``libvbg_sm100a.so`` (``include/vbg.h``).

Mirrors the reference's operator interface for the hot path
(reference model/ViBERTgrid_net.py):
  * constructor kwargs                       :128-159
  * ``forward(image, seg_indices, segment_classes, coors, corpus, mask)``   :501-544
  * ``inference(image, seg_indices, coors, corpus, mask)``                  :470-499
  * ``train()/eval()`` overrides and their ``work_mode`` quirk              :462-468
  * state-dict key layout, incl. the aliased
    ``BERTgrid_generator.model.*`` copy of ``bert_model.*``                 :358-362

There is no CPU or PyTorch-eager fallback: every stage below calls a kernel in
``csrc/``; if the library is missing, ``ops`` raises at import/first use.
"""
from __future__ import annotations

import json
import os
import warnings
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import params as P

_BERT_NAMES = {
    "private_bert-base-uncased": 768, "bert-base-uncased": 768, "bert-base-cased": 768,
    "roberta-base": 768, "bert-base-chinese": 768, "hfl/chinese-bert-wwm-ext": 768,
    "hfl/chinese-bert-wwm": 768,
}
_BERT_DEFAULT_CFG = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                         intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2,
                         hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12, pad_token_id=0)
# roberta-base's published hyper-parameters where they differ from bert-base
_ROBERTA_DEFAULT_CFG = dict(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1)


def _bert_config(name: str) -> dict:
    """Hyper-parameters of ``name``: a local directory's config.json first (works
    offline), then HuggingFace ``AutoConfig``, then bert-base defaults."""
    cfg = dict(_BERT_DEFAULT_CFG)
    if "roberta-" in name:
        cfg.update(_ROBERTA_DEFAULT_CFG)
    path = os.path.join(name, "config.json")
    src = None
    if os.path.isfile(path):
        with open(path) as f:
            src = json.load(f)
    else:
        try:
            from transformers import AutoConfig
            src = AutoConfig.from_pretrained(name).to_dict()
        except Exception as e:  # offline and not cached
            warnings.warn(f"[vibertgrid_b200] no config for {name!r} ({type(e).__name__}); using bert-base defaults")
    if src:
        for k in cfg:
            if k in src and src[k] is not None:
                cfg[k] = src[k]
    # RobertaModel numbers positions from padding_idx + 1 and skips <pad> ids (HF create_position_ids_from_input_ids)
    cfg["roberta"] = "roberta-" in name
    return cfg


def _load_tokenizer(name, tokenizer):
    if tokenizer is not None:
        # the reference accepts only an instance of the model family's tokenizer class (model/ViBERTgrid_net.py:235-252)
        try:
            from transformers import BertTokenizer, RobertaTokenizer
            cls = RobertaTokenizer if "roberta-" in name else BertTokenizer
        except Exception:       # pragma: no cover - transformers missing: nothing to check against
            return tokenizer
        if not isinstance(tokenizer, cls):
            raise ValueError(f"invalid value of parameter tokenizer, must be None or callable {cls.__name__}")
        return tokenizer
    try:
        from transformers import BertTokenizer, RobertaTokenizer
        cls = RobertaTokenizer if "roberta-" in name else BertTokenizer
        return cls.from_pretrained(name)
    except Exception:
        return None


class _BERTgridGeneratorParams(P._NoForward):
    """Holds the alias ``model`` -> ``bert_model`` so ``state_dict()`` emits both key sets."""

    def __init__(self, bert, grid_mode, stride):
        super().__init__()
        self.model = bert
        self.grid_mode, self.stride = grid_mode, stride


class ViBERTgridNet(nn.Module):
    """B200-native ViBERTgrid joint forward behind the reference's module interface."""

    def __init__(self, num_classes, image_mean: Any, image_std: Any, image_min_size: Any, image_max_size,
                 test_image_min_size=512, bert_model: str = "bert-base-uncased", tokenizer: Any = None,
                 backbone: str = "resnet_18_fpn", grid_mode: str = "mean", early_fusion_downsampling_ratio=8,
                 roi_shape=7, p_fuse_downsampling_ratio=4, late_fusion_fuse_embedding_channel=1024,
                 loss_weights: Any = None, num_hard_positive_main_1=-1, num_hard_negative_main_1=-1,
                 num_hard_positive_main_2=-1, num_hard_negative_main_2=-1, loss_aux_sample_list: List = None,
                 num_hard_positive_aux=-1, num_hard_negative_aux=-1, loss_control_lambda: float = 1,
                 add_pos_neg: bool = True, classifier_mode: str = "full", tag_to_idx: Dict = None,
                 ohem_random: bool = False, layer_mode: str = "single", work_mode: str = "train") -> None:
        super().__init__()
        assert work_mode in ("train", "eval", "inference"), \
            f"mode must be 'train' 'eval' or 'inference', {work_mode} given"
        self.work_mode = work_mode
        self.num_classes = num_classes
        self.num_tokens = len(tag_to_idx) if tag_to_idx is not None else num_classes

        def _triple(v, what):
            assert isinstance(v, (float, list)), f"{what} must be float or list of float, {type(v)} given"
            if isinstance(v, float):
                return [v] * 3
            if len(v) != 3:
                raise ValueError(f"{what} must contain 3 three values, {len(v)} given")
            return list(v)

        self.image_mean, self.image_std = _triple(image_mean, "image_mean"), _triple(image_std, "image_std")
        self.test_image_min_size = test_image_min_size
        assert isinstance(image_min_size, (int, tuple, list)), \
            f"image_min_size must be int, Tuple or List, {type(image_min_size)} given"
        self.image_min_size = list(image_min_size) if not isinstance(image_min_size, int) else [image_min_size]
        assert isinstance(image_max_size, int), f"image_max_size must be int, {type(image_max_size)} given"
        self.image_max_size = image_max_size

        assert bert_model in _BERT_NAMES, \
            f"the given bert model {bert_model} does not exists, see attribute bert_model_list for all bert_models"
        self.bert_model_list = dict(_BERT_NAMES)
        self.bert_hidden_size = _BERT_NAMES[bert_model]
        self.tokenizer = _load_tokenizer(bert_model, tokenizer)
        self.bert_cfg = _bert_config(bert_model)
        self.bert_model = P.BertParams(**self.bert_cfg)
        self.bert_hidden_dropout = float(self.bert_cfg["hidden_dropout_prob"])      # training-mode forward only
        self.bert_attn_dropout = float(self.bert_cfg.get("attention_probs_dropout_prob", 0.0))
        if work_mode in ("train", "inference"):
            print("loading pretrained")
            self._load_pretrained_bert(bert_model)
        else:
            print("in evaluation mode, no pretrained will be loaded")

        assert backbone in P.BACKBONES, \
            f"the given backbone {backbone} does not exists, see attribute backbone_list for all backbones"
        self.backbone_list = list(P.BACKBONES)
        self.backbone_name = backbone
        self.backbone = P.BACKBONES[backbone](self.bert_hidden_size)
        if backbone.endswith("_pretrained") and not self.backbone.try_load_hub_weights():
            # the reference calls torchvision's resnetXX(pretrained=True) (model/ResNetFPN_ViBERTgrid.py:521,524) and fails when
            # the weights cannot be fetched; training from a silent random init is never what a drop-in user wants
            if work_mode in ("train", "inference") and os.environ.get("VBG_ALLOW_RANDOM_INIT") != "1":
                raise RuntimeError(f"[vibertgrid_b200] torchvision hub weights for {backbone!r} are not cached "
                                   "($TORCH_HOME/hub/checkpoints) and cannot be downloaded; set VBG_ALLOW_RANDOM_INIT=1 to "
                                   "train from a random initialisation")
            warnings.warn("[vibertgrid_b200] torchvision hub weights not cached; backbone keeps random init")
        self.p_fuse_channel = 256

        assert grid_mode in ("mean", "first"), f"grid_mode should be 'mean' or 'first', {grid_mode} were given"
        self.grid_mode = grid_mode
        self.early_fusion_downsampling_ratio = early_fusion_downsampling_ratio
        self.roi_shape = roi_shape
        self.p_fuse_downsampling_ratio = p_fuse_downsampling_ratio
        self.late_fusion_fuse_embedding_channel = late_fusion_fuse_embedding_channel
        self.loss_control_lambda = None if work_mode == "inference" else loss_control_lambda
        if loss_weights is None or work_mode == "inference":
            self.loss_weights = None
        elif isinstance(loss_weights, list):
            self.loss_weights = torch.tensor(loss_weights)
        elif isinstance(loss_weights, torch.Tensor):
            self.loss_weights = loss_weights
        else:
            raise TypeError(f"loss_weights must be None, List or torch.Tensor, {type(loss_weights)} given")
        assert classifier_mode in ("full", "simp", "crf"), "invalid classifier mode, must be 'full', 'simp' or 'crf'"
        self.classifier_mode = classifier_mode
        self.layer_mode = layer_mode
        self.add_pos_neg = add_pos_neg

        self.BERTgrid_generator = _BERTgridGeneratorParams(self.bert_model, grid_mode, early_fusion_downsampling_ratio)
        self.late_fusion_net = P.LateFusionParams(self.bert_hidden_size, self.p_fuse_channel, roi_shape)
        c = late_fusion_fuse_embedding_channel
        if classifier_mode == "full":
            self.field_type_classification_head = P.FullHeadParams(self.num_tokens, c, layer_mode)
        elif classifier_mode == "simp":
            assert layer_mode in ("single", "multi"), f"layer_mode must be single or multi, {layer_mode} given"
            self.field_type_classification_head = P.SimpHeadParams(self.num_tokens, c, work_mode)
        else:
            assert tag_to_idx is not None, "tag_to_idx cannot be None in crf mode"
            assert max(tag_to_idx.values()) == len(tag_to_idx) - 1, "invalid tag_to_idx format"
            n = len(tag_to_idx)
            tag_to_idx["<START>"], tag_to_idx["<STOP>"] = n, n + 1      # mutated in place like the reference
            self.tag_to_idx = tag_to_idx
            self.field_type_classification_head = P.CRFHeadParams(n, c, layer_mode)
        if work_mode == "inference":
            self.semantic_segmentation_head = None
        else:
            self.semantic_segmentation_head = P.SegHeadParams(self.p_fuse_channel, self.num_tokens,
                                                              simplified=(classifier_mode == "simp"))
        if self.loss_weights is not None:
            # the reference's weighted loss modules are nn.CrossEntropyLoss / nn.BCEWithLogitsLoss subclasses: their class-weight
            # tensor is a registered buffer, so it shows up in state_dict() (pipeline/custom_loss.py:13-20,109-117,298-304;
            # semantic_segmentation_head.py:140-149,270-277; field_type_classification_head.py:262-283,486-500)
            seg, head = self.semantic_segmentation_head, self.field_type_classification_head
            if classifier_mode == "simp":
                seg.add_module("aux_loss_2", P.LossWeightParams(self.loss_weights))
                head.add_module("field_type_classification_loss", P.LossWeightParams(self.loss_weights))
            else:
                for i in range(self.num_tokens - 1):
                    seg.add_module(f"aux_loss_2_{i}", P.LossWeightParams(self.loss_weights))
                    if classifier_mode == "full":
                        head.add_module(f"field_type_classification_loss_{i}", P.LossWeightParams(self.loss_weights))
        self.loss_cfg = dict(main_1=(num_hard_positive_main_1, num_hard_negative_main_1),
                             main_2=(num_hard_positive_main_2, num_hard_negative_main_2),
                             aux_sample_list=loss_aux_sample_list,
                             aux=(num_hard_positive_aux, num_hard_negative_aux), random=ohem_random)
        self._engine = None
        self._train_engine = None

    # ------------------------------------------------------------------ construction helpers
    def _load_pretrained_bert(self, name):
        try:
            from transformers import BertModel, RobertaModel
            hf = (RobertaModel if "roberta-" in name else BertModel).from_pretrained(name)
            missing = self.bert_model.load_state_dict(hf.state_dict(), strict=False)
            if missing.missing_keys:
                warnings.warn(f"[vibertgrid_b200] pretrained BERT lacks keys: {missing.missing_keys[:4]}...")
        except Exception as e:
            # the reference's BertModel.from_pretrained(...) raises here (model/ViBERTgrid_net.py:232-253); so do we, unless the
            # caller opted into a random initialisation explicitly
            if os.environ.get("VBG_ALLOW_RANDOM_INIT") != "1":
                raise
            warnings.warn(f"[vibertgrid_b200] could not load pretrained {name!r} ({type(e).__name__}: {e}); "
                          "keeping BERT-style random init (VBG_ALLOW_RANDOM_INIT=1)")

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # checkpoints from transformers<4.31 carry a persistent position_ids buffer
        sd = {k: v for k, v in state_dict.items() if not k.endswith("embeddings.position_ids")}
        out = super().load_state_dict(sd, strict=strict, **kw)
        if self._engine is not None:
            self._engine.invalidate()
        return out

    # ------------------------------------------------------------------ mode quirk
    def train(self, mode: bool = True):
        self.work_mode = "train"
        return super().train(mode)

    def eval(self):
        self.work_mode = "eval"
        return super().eval()

    # ------------------------------------------------------------------ the hot path
    def _get_engine(self):
        if self._engine is None:
            from .engine import ForwardEngine
            self._engine = ForwardEngine(self)
        return self._engine

    def inference(self, image, seg_indices, coors, corpus, mask):
        # the reference's deployment path hands ``seg_indices`` over as HOST tensors (deployment/inference_preporcessing.py:184:
        # only image / coors / corpus / mask are moved to the device; its model reads them with .item() in Python loops)
        dev = corpus.device
        if dev.type == "cuda" and any(s.device != dev for s in seg_indices):
            seg_indices = tuple(s.to(dev, non_blocking=True) for s in seg_indices)
        out = self._get_engine().run(image, seg_indices, None, coors, corpus, mask, want_seg=False, crf_one_sequence=True)
        return out["pred_label"].clone() if out.get("static") else out["pred_label"]

    def forward(self, image, seg_indices, segment_classes, coors, corpus, mask):
        if self.training:
            # model/ViBERTgrid_net.py:541: in training the module returns the loss alone (a differentiable 0-d tensor)
            if self._train_engine is None:
                from .train_engine import TrainEngine
                self._train_engine = TrainEngine(self)
            return self._train_engine.loss(image, seg_indices, segment_classes, coors, corpus, mask)
        eng = self._get_engine()
        out = eng.run(image, seg_indices, segment_classes, coors, corpus, mask, want_seg=True)
        from . import losses
        loss_aux = losses.aux_loss(self, out)
        loss_c = losses.main_loss(self, out)
        total_loss = loss_c + self.loss_control_lambda * loss_aux
        self.last_intermediates = out
        ret = (out["pred_mask"], out["pred_ss"], out["gt_label"], out["pred_label"])
        if out.get("static"):          # CUDA-graph buffers are rewritten by the next call: hand out copies
            ret = tuple(t.clone() for t in ret)
        return (total_loss,) + ret
