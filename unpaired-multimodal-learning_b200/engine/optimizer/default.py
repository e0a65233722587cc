"""Hyper-parameter presets, keyed like the reference's ``HYPER_DICT``
(engine/optimizer/default.py:1-60).  List-valued entries are swept (cartesian product) by
``finetune.sweep``; scalars are fixed."""


def _preset(lr, batch_size, patience, learnable_temp, weight_decay=(0.0, 0.01, 0.001)):
    return {
        "optim": "adamw",
        "lr": list(lr),
        "weight_decay": list(weight_decay),
        "lr_scheduler": "cosine",
        "batch_size": list(batch_size),
        "max_iter": [12800],
        "warmup_iter": 50,
        "warmup_type": "linear",
        "warmup_min_lr": 1e-5,
        "dropout": [0.0],
        "learnable_temp": [learnable_temp],
        "patience": [patience],
    }


HYPER_DICT = {
    # full finetuning experiments
    "full_ds_full_model_finetune": _preset(lr=[5e-05], batch_size=[64], patience=10, learnable_temp=False),
    # linear probe on CLIP encoders
    "clip_linear": _preset(lr=[0.001, 0.0001], batch_size=[32], patience=5, learnable_temp=False),
    # linear probe on unimodal vision + language encoders
    "linear": _preset(lr=[0.001, 0.0001], batch_size=[8, 32], patience=10, learnable_temp=True),
    "audio": _preset(lr=[0.1, 0.01, 0.001, 0.0001], batch_size=[8], patience=5, learnable_temp=False,
                     weight_decay=(0.0, 0.01, 0.0001)),
}
