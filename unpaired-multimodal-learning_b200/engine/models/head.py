"""Shared-head models over pre-extracted features (L2 of the hot path).

Mirrors ``engine/models/head.py`` of the reference: ``UML`` (:39-98) = optional linear adapter
``img_proj`` + shared ``head`` + per-modality logit scales; ``UMLClip`` (:101-141) = shared head with
CLIP's fixed ``exp(logit_scale)``.  State-dict keys are the reference's (``head.weight``,
``img_proj.weight``, ``img_scale``, ``txt_scale``); backbone keys do not exist because the frozen
backbone was applied offline by ``features.py`` - load reference checkpoints with ``strict=False``.

The modules are parameter containers: the training step never runs autograd, it hands the weights to
the CUDA kernels (``engine/trainer.py``).  ``forward`` is provided for API compatibility and runs the
same kernels without building a graph.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from ... import ops

# feature widths of the encoders the reference's config admits (engine/config/__init__.py:74-111)
# plus the two synthetic-shape names used by BASELINE.json
CLIP_EMBED_DIM = {"RN50": 1024, "RN101": 512, "ViT-B/32": 512, "ViT-B/16": 512, "ViT-L/14": 768}
VISION_NUM_FEATURES = {
    "vit_base_patch16_224_dino": 768, "vit_base_patch8_224_dino": 768,
    "vit_small_patch14_dinov2.lvd142m": 384, "vit_base_patch14_dinov2.lvd142m": 768,
    "vit_large_patch14_dinov2.lvd142m": 1024, "vit_giant_patch14_dinov2.lvd142m": 1536,
}
LANGUAGE_HIDDEN = {
    "bert-base-uncased": 768, "bert-large-uncased": 1024, "roberta-base": 768, "roberta-large": 1024,
    "openlm-research/open_llama_3b_v2": 3200, "meta-llama/Llama-2-7b-chat-hf": 4096, "gpt2": 768,
    "gpt2-medium": 1024, "gpt2-large": 1280, "mistralai/Mistral-7B-v0.1": 4096, "bigscience/bloom-1b1": 1536,
}


def _width(name_or_dim, table, what):
    if isinstance(name_or_dim, int):
        return name_or_dim
    s = str(name_or_dim)
    if ":" in s and s.split(":")[-1].isdigit():  # "bank:768" style explicit width
        return int(s.split(":")[-1])
    if s in table:
        return table[s]
    raise ValueError(f"unknown {what} {name_or_dim!r}: pass the feature width as an int or 'name:<dim>'")


def get_zero_shot_weights(text_dataset, num_classes, in_features, device="cuda"):
    """Class-mean text features, L2-normalised per row; classes without text rows stay all-zero
    (reference head.py:22-37 - ``F.normalize`` clamps the norm at 1e-12 so 0 stays 0)."""
    feats = getattr(text_dataset, "input_tensor", None)
    labels = getattr(text_dataset, "label_tensor", None)
    if feats is None:  # a FeatureBank
        feats, labels = text_dataset.features, text_dataset.labels
    # One-off initialisation, done on the host: the CPU index_add_ accumulates rows in order, exactly like the
    # reference's Python loop, and - unlike the atomic CUDA version - gives the same bits on every rank of a
    # data-parallel run (replicas that start 1 ulp apart never re-converge).
    feats = feats.detach().to(device="cpu", dtype=torch.float32)
    labels = labels.detach().to(device="cpu", dtype=torch.int64)
    with torch.no_grad():
        sums = torch.zeros(num_classes, in_features).index_add_(0, labels, feats)
        counts = torch.bincount(labels, minlength=num_classes).clamp_min(1).unsqueeze(1)
        w = sums / counts
        w = w / w.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return w.cpu()


class _HeadBase(torch.nn.Module):
    precision = "auto"  # "fp32" (SIMT, exact) | "bf16" (tcgen05) | "auto" (by batch size)

    def _linear(self, x: torch.Tensor, weight: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
        out = torch.empty((x.shape[0], weight.shape[0]), device=x.device, dtype=torch.float32)
        ops.gemm_nt(x.contiguous(), weight.detach(), out, alpha=alpha)
        return out

    def zero_shot_init(self, zeroshot_dataset):
        print("=> Initializing head with zero-shot weights")
        dev = self.head.weight.device
        w = get_zero_shot_weights(zeroshot_dataset, self.num_classes, self.shared_dim,
                                  device=dev if dev.type == "cuda" else "cpu")
        self.head.weight.data = w.to(dev)


class UML(_HeadBase):
    """``UML(vision_model, text_indim, num_classes, bias=False, learnable_temp=False, freeze_backbone=False)``

    ``vision_model`` names the (offline) image encoder only to fix the image feature width; an int or
    ``'name:<dim>'`` is accepted for encoders outside the table."""

    def __init__(self, vision_model, text_indim, num_classes, bias=False, learnable_temp=False, freeze_backbone=False):
        super().__init__()
        if bias:
            raise NotImplementedError("the reference always builds its heads with bias=False (finetune.py:338-346)")
        self.num_classes = num_classes
        self.img_indim = _width(vision_model, VISION_NUM_FEATURES, "vision model")
        self.img_proj = None
        self.shared_dim = self.img_indim
        if text_indim > 0:
            self.img_proj = torch.nn.Linear(self.img_indim, text_indim, bias=False)
            self.shared_dim = text_indim
        self.head = torch.nn.Linear(self.shared_dim, num_classes, bias=False)
        self.learnable_temp = bool(learnable_temp)
        self.img_scale = torch.nn.Parameter(torch.tensor(1.0)) if learnable_temp else torch.tensor(1.0)
        self.txt_scale = torch.nn.Parameter(torch.tensor(1.0)) if learnable_temp else torch.tensor(1.0)
        for p in self.parameters():
            p.requires_grad_(False)  # gradients are produced by the CUDA kernels, not autograd
        total = sum(p.numel() for p in self.parameters())
        print(f"=> Model trainable params after init: {total}/{total}")

    def scales(self):
        return self.img_scale, self.txt_scale

    def extract_raw_features(self, images):
        return images

    def extract_features(self, images):
        return self._linear(images, self.img_proj.weight) if self.img_proj is not None else images

    def forward(self, images, text_features=None):
        z = self.extract_features(images)
        img_logits = self._linear(z, self.head.weight, float(self.img_scale))
        if text_features is None:
            return img_logits, None
        return img_logits, self._linear(text_features, self.head.weight, float(self.txt_scale))


class UMLClip(_HeadBase):
    """``UMLClip(clip_encoder, num_classes, logit_scale_init=log(1/0.07), ...)`` - shared linear head on
    CLIP image/text features, logits multiplied by the fixed ``exp(logit_scale)`` (head.py:131-137)."""

    def __init__(self, clip_encoder, num_classes, logit_scale_init=math.log(1 / 0.07), bias=False,
                 learnable_temp=False, freeze_backbone=False):
        super().__init__()
        if bias:
            raise NotImplementedError("the reference always builds its heads with bias=False")
        self.num_classes = num_classes
        self.img_proj = None
        self.shared_dim = self.img_indim = _width(clip_encoder, CLIP_EMBED_DIM, "CLIP encoder")
        self.head = torch.nn.Linear(self.shared_dim, num_classes, bias=False)
        self.logit_scale = torch.tensor(float(logit_scale_init))  # fixed, not a Parameter, not in the state dict
        self.learnable_temp = False
        for p in self.parameters():
            p.requires_grad_(False)
        total = sum(p.numel() for p in self.parameters())
        print(f"=> CLIP-model trainable params after init: {total}/{total}")

    def scales(self):
        s = self.logit_scale.exp()
        return s, s

    def extract_raw_features(self, images):
        return images

    def extract_features(self, images):  # the reference forgot this one (finetune.py:183 needs it)
        return images

    def forward(self, images, text_features=None):
        s = float(self.logit_scale.exp())
        img_logits = self._linear(images, self.head.weight, s)
        if text_features is None:
            return img_logits, None
        return img_logits, self._linear(text_features, self.head.weight, s)
