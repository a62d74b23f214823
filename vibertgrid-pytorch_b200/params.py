"""Parameter containers with the reference's state-dict layout.

These ``nn.Module`` trees own the weights only; none of their ``forward``
methods is ever called -- the arithmetic is done by the sm_100a kernels in
``csrc/`` through the C-ABI (``include/vbg.h``).  Their job is to make
``state_dict()`` / ``load_state_dict()`` / ``named_parameters()`` of the
drop-in ``ViBERTgridNet`` emit exactly the keys the reference emits
(SURVEY.md Appendix D), so reference checkpoints round-trip and the
reference's optimiser split on the substring ``"bert_model"``
(reference train_SROIE.py:217-221) keeps working.

Layout sources (names only, no code taken):
  * BERT encoder      -- HuggingFace ``BertModel`` key names as used at
                         reference model/ViBERTgrid_net.py:253,268
  * plain / D ResNet  -- reference model/ResNetFPN_ViBERTgrid.py:106-184 (block),
                         :187-269 (D block), :272-321 (early fusion), :324-464
  * pretrained ResNet -- torchvision ``resnet18/34`` names under ``resnet.``,
                         reference model/ResNetFPN_ViBERTgrid.py:511-610
  * late fusion/heads -- reference model/field_type_classification_head.py:26-190,
                         :410-519, :591-653, :193-296
  * seg head          -- reference model/semantic_segmentation_head.py:23-64,:100-159,:236-286
"""
from __future__ import annotations

import torch
import torch.nn as nn

BN_EPS = 1e-5
LN_EPS = 1e-12


def _conv(cin, cout, k, stride, pad, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=bias)


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; the forward runs in the "
            "sm_100a kernels (vibertgrid_pytorch_b200.net.ViBERTgridNet.forward)"
        )


# --------------------------------------------------------------------------- BERT
class BertEmbeddingsParams(_NoForward):
    def __init__(self, vocab, hidden, max_pos, type_vocab, eps=LN_EPS, pad_id=0, pad_positions=False):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, hidden, padding_idx=pad_id)
        # RobertaEmbeddings gives the position table a padding row as well
        self.position_embeddings = nn.Embedding(max_pos, hidden, padding_idx=pad_id if pad_positions else None)
        self.token_type_embeddings = nn.Embedding(type_vocab, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)


class _SelfParams(_NoForward):
    def __init__(self, hidden):
        super().__init__()
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(hidden, hidden)
        self.value = nn.Linear(hidden, hidden)


class _DenseLN(_NoForward):
    def __init__(self, cin, cout, eps=LN_EPS):
        super().__init__()
        self.dense = nn.Linear(cin, cout)
        self.LayerNorm = nn.LayerNorm(cout, eps=eps)


class _Dense(_NoForward):
    def __init__(self, cin, cout):
        super().__init__()
        self.dense = nn.Linear(cin, cout)


class _AttnParams(_NoForward):
    def __init__(self, hidden, eps=LN_EPS):
        super().__init__()
        self.self = _SelfParams(hidden)
        self.output = _DenseLN(hidden, hidden, eps)


class BertLayerParams(_NoForward):
    def __init__(self, hidden, inter, eps=LN_EPS):
        super().__init__()
        self.attention = _AttnParams(hidden, eps)
        self.intermediate = _Dense(hidden, inter)
        self.output = _DenseLN(inter, hidden, eps)


class _Encoder(_NoForward):
    def __init__(self, n_layers, hidden, inter, eps=LN_EPS):
        super().__init__()
        self.layer = nn.ModuleList([BertLayerParams(hidden, inter, eps) for _ in range(n_layers)])


class BertParams(_NoForward):
    """Weights of a post-LN BERT encoder under HuggingFace key names.

    ``pooler.dense`` exists (and never receives a gradient)
    because the reference's state dict carries it.
    """

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072,
                 max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=LN_EPS, pad_token_id=0, roberta=False,
                 **_unused):
        super().__init__()
        self.cfg = dict(vocab_size=vocab_size, hidden_size=hidden_size,
                        num_hidden_layers=num_hidden_layers,
                        num_attention_heads=num_attention_heads,
                        intermediate_size=intermediate_size,
                        max_position_embeddings=max_position_embeddings,
                        type_vocab_size=type_vocab_size, layer_norm_eps=layer_norm_eps, pad_token_id=pad_token_id,
                        roberta=bool(roberta))
        self.embeddings = BertEmbeddingsParams(vocab_size, hidden_size, max_position_embeddings, type_vocab_size,
                                               layer_norm_eps, pad_token_id, pad_positions=bool(roberta))
        self.encoder = _Encoder(num_hidden_layers, hidden_size, intermediate_size, layer_norm_eps)
        self.pooler = _Dense(hidden_size, hidden_size)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        # BERT's published initialisation: N(0, 0.02), LN = (1, 0)
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, 0.0, 0.02)
            if isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)
            elif m.padding_idx is not None:
                with torch.no_grad():
                    m.weight[m.padding_idx].zero_()


# --------------------------------------------------------------------------- ResNet-FPN
class ResBlockParams(_NoForward):
    """conv_1/bn_1/conv_2/bn_2 (+ conv_shortcut) -- plain and "D" variants."""

    def __init__(self, cin, cout, downsample=False, d_variant=False):
        super().__init__()
        self.downsample, self.d_variant = downsample, d_variant
        if downsample:
            self.conv_1 = _conv(cin, cout, 3, 2, 1)
            if d_variant:
                self.conv_shortcut = nn.Sequential(nn.AvgPool2d(2, 2), _conv(cin, cout, 1, 1, 0),
                                                   nn.BatchNorm2d(cout))
            else:
                self.conv_shortcut = nn.Sequential(_conv(cin, cout, 1, 2, 0), nn.BatchNorm2d(cout))
        else:
            # quirk kept: a non-downsampling block is always cout -> cout
            self.conv_1 = _conv(cout, cout, 3, 1, 1)
            self.conv_shortcut = nn.Identity()
        self.bn_1 = nn.BatchNorm2d(cout)
        self.conv_2 = _conv(cout, cout, 3, 1, 1)
        self.bn_2 = nn.BatchNorm2d(cout)


class EarlyFusionStageParams(_NoForward):
    def __init__(self, cin, cout, n_blocks, grid_channel, d_variant):
        super().__init__()
        self.block_1 = ResBlockParams(cin, cout, True, d_variant)
        self.early_fusion = _conv(cout + grid_channel, cout, 1, 1, 0, bias=True)
        self.layers = nn.Sequential(*[ResBlockParams(cin, cout, False, d_variant)
                                      for _ in range(n_blocks - 1)])


def _stage(cin, cout, n, downsample, d_variant):
    return nn.Sequential(*[ResBlockParams(cin if i == 0 else cout, cout,
                                          downsample if i == 0 else False, d_variant)
                           for i in range(n)])


class _FPNParams(_NoForward):
    """1x1 laterals, 3x3 merges, 1x1 fuse: no bias / BN / activation."""

    def _make_fpn(self, p=256, f=256):
        self.conv_6_x = _conv(512, p, 1, 1, 0)
        self.skip_1 = _conv(256, p, 1, 1, 0)
        self.merge_1 = _conv(p, p, 3, 1, 1)
        self.skip_2 = _conv(128, p, 1, 1, 0)
        self.merge_2 = _conv(p, p, 3, 1, 1)
        self.skip_3 = _conv(64, p, 1, 1, 0)
        self.merge_3 = _conv(p, p, 3, 1, 1)
        self.fuse = _conv(4 * p, f, 1, 1, 0)


class ResNetFPNParams(_FPNParams):
    def __init__(self, sizes, grid_channel, d_variant=False):
        super().__init__()
        self.sizes, self.d_variant, self.pretrained_layout = list(sizes), d_variant, False
        self.conv_1 = nn.Sequential(_conv(3, 64, 7, 2, 3), nn.BatchNorm2d(64), nn.ReLU(inplace=True))
        self.conv_2_x = _stage(64, 64, sizes[0], False, d_variant)
        self.conv_3_x = EarlyFusionStageParams(64, 128, sizes[1], grid_channel, d_variant)
        self.conv_4_x = _stage(128, 256, sizes[2], True, d_variant)
        self.conv_5_x = _stage(256, 512, sizes[3], True, d_variant)
        self._make_fpn()


class _TVBlock(_NoForward):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride, 1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = _conv(cout, cout, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(_conv(cin, cout, 1, stride, 0), nn.BatchNorm2d(cout))


class _TVResNet(_NoForward):
    """torchvision resnet18/34 key layout, including the unused ``fc``."""

    def __init__(self, sizes):
        super().__init__()
        self.conv1 = _conv(3, 64, 7, 2, 3)
        self.bn1 = nn.BatchNorm2d(64)
        chans = [64, 128, 256, 512]
        cin = 64
        for li, (c, n) in enumerate(zip(chans, sizes)):
            blocks = []
            for i in range(n):
                blocks.append(_TVBlock(cin, c, 2 if (i == 0 and li > 0) else 1))
                cin = c
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        self.fc = nn.Linear(512, 1000)


class ResNetFPNPretrainedParams(_FPNParams):
    def __init__(self, resnet_type, grid_channel):
        super().__init__()
        sizes = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3]}[resnet_type]
        self.sizes, self.d_variant, self.pretrained_layout = sizes, False, True
        self.resnet_type = resnet_type
        self.resnet = _TVResNet(sizes)
        self.early_fusion = _conv(grid_channel + 128, 128, 1, 1, 0, bias=False)
        self._make_fpn()

    def try_load_hub_weights(self):
        """Mirror of ``resnetXX(pretrained=True)`` without a download: use the
        torch-hub cache file when it is present (SURVEY 8c bridge (2))."""
        import os
        fname = {"resnet18": "resnet18-f37072fd.pth", "resnet34": "resnet34-b627a593.pth"}[self.resnet_type]
        path = os.path.join(torch.hub.get_dir(), "checkpoints", fname)
        if os.path.isfile(path):
            self.resnet.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
            return True
        return False


BACKBONES = {
    "resnet_18_fpn": lambda g: ResNetFPNParams([2, 2, 2, 2], g, False),
    "resnet_34_fpn": lambda g: ResNetFPNParams([3, 4, 6, 3], g, False),
    "resnet_18_D_fpn": lambda g: ResNetFPNParams([2, 2, 2, 2], g, True),
    "resnet_34_D_fpn": lambda g: ResNetFPNParams([3, 4, 6, 3], g, True),
    "resnet_18_fpn_pretrained": lambda g: ResNetFPNPretrainedParams("resnet18", g),
    "resnet_34_fpn_pretrained": lambda g: ResNetFPNPretrainedParams("resnet34", g),
}


# --------------------------------------------------------------------------- heads
class _Lin(_NoForward):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear = nn.Linear(cin, cout)


class _MLP(_NoForward):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cin // 2)
        self.linear_2 = nn.Linear(cin // 2, cout)


class _Binary(_NoForward):
    def __init__(self, cin, layer_mode):
        super().__init__()
        self.layer = _Lin(cin, 1) if layer_mode == "single" else _MLP(cin, 1)


class ROIEmbeddingParams(_NoForward):
    def __init__(self, c, roi):
        super().__init__()
        self.conv_1 = _conv(c, c, 3, 1, 1)
        self.bn_1 = nn.BatchNorm2d(c)
        self.conv_2 = _conv(c, c, 3, 1, 1)
        self.bn_2 = nn.BatchNorm2d(c)
        self.linear = nn.Linear(c * roi * roi, 1024)


class LateFusionParams(_NoForward):
    def __init__(self, bert_hidden, roi_channel, roi_shape):
        super().__init__()
        self.ROI_embedding_net = ROIEmbeddingParams(roi_channel, roi_shape)
        self.fuse_embedding_net = _Lin(bert_hidden + 1024, 1024)


class SimpHeadParams(_NoForward):
    """``simp`` head: always the two-layer MLP (reference typo)."""

    def __init__(self, num_classes, c, work_mode):
        super().__init__()
        if work_mode != "inference":
            self.pos_neg_classification_net = _MLP(c, 2)
        self.category_classification_net = _MLP(c, num_classes)


class FullHeadParams(_NoForward):
    def __init__(self, num_classes, c, layer_mode):
        super().__init__()
        self.pos_neg_classification_net = _Binary(c, layer_mode)
        for i in range(num_classes - 1):
            self.add_module(f"category_classification_net_{i}", _Binary(c, layer_mode))


class CRFParams(_NoForward):
    def __init__(self, n_tags, start, stop):
        super().__init__()
        self.transitions = nn.Parameter(torch.randn(n_tags, n_tags))
        with torch.no_grad():
            self.transitions[start, :] = -10000
            self.transitions[:, stop] = -10000


class CRFHeadParams(_NoForward):
    def __init__(self, num_classes, c, layer_mode):
        super().__init__()
        t = num_classes + 2
        self.category_classification_net = _Lin(c, t) if layer_mode == "single" else _MLP(c, t)
        self.crf_layer = CRFParams(t, num_classes, num_classes + 1)


class SegEncoderParams(_NoForward):
    def __init__(self, c, num_classes):
        super().__init__()
        self.conv_1 = _conv(c, c, 3, 1, 1)
        self.bn_1 = nn.BatchNorm2d(c)
        self.conv_2 = _conv(c, c, 3, 1, 1)
        self.bn_2 = nn.BatchNorm2d(c)
        self.conv_3_1 = _conv(c, 3, 1, 1, 0, bias=True)
        self.conv_3_2 = _conv(c, num_classes, 1, 1, 0, bias=True)


class LossWeightParams(_NoForward):
    """State-dict stand-in of a weighted loss module of the reference (``<loss>.weight`` buffer)."""

    def __init__(self, weight):
        super().__init__()
        self.register_buffer("weight", weight)


class _SegBinary(_NoForward):
    def __init__(self, cin):
        super().__init__()
        self.conv1 = _conv(cin, 1, 1, 1, 0, bias=True)


class SegHeadParams(_NoForward):
    """``simplified=True`` -> ``semantic_segmentation_encoder.*``; else ``ss_encoder.*``
    plus ``ss_binary_classifier_{i}.conv1``."""

    def __init__(self, c, num_classes, simplified):
        super().__init__()
        self.simplified = simplified
        enc = SegEncoderParams(c, num_classes)
        if simplified:
            self.semantic_segmentation_encoder = enc
        else:
            self.ss_encoder = enc
            for i in range(num_classes - 1):
                self.add_module(f"ss_binary_classifier_{i}", _SegBinary(num_classes))

    @property
    def encoder(self):
        return self.semantic_segmentation_encoder if self.simplified else self.ss_encoder
