"""Synthetic inputs with the reference's batch contract (model level; v7.00/src/dataset.py:58-61
for the image-first layout): input_ids / labels / images / sample_id."""
from __future__ import annotations

import torch

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = 65535


def make_batch(B, T, n_img_tok, image_size, seed=0, vocab=65536, device="cpu", img_dtype=torch.float32,
               human_tokens=40):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 1, (B, T), generator=g)
    labels = ids.clone()
    for b in range(B):
        start = int(torch.randint(0, 9, (1,), generator=g))
        ids[b, start:start + n_img_tok] = IMAGE_TOKEN_INDEX
        labels[b, : start + n_img_tok + human_tokens] = IGNORE_INDEX
    images = torch.randn(B, 3, image_size, image_size, generator=g)
    return {"input_ids": ids.to(device), "labels": labels.to(device), "images": images.to(device=device, dtype=img_dtype),
            "sample_id": [str(i) for i in range(B)]}
