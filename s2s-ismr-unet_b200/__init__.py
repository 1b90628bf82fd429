"""s2s-ismr-unet_b200 — B200-native U-Net hot path (host side).

    csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/s2s_unet.h)
    build.py     nvcc build of lib/libs2s_unet.so
    _lib.py      ctypes binding generated from the public header
    runtime.py   streams / device + pinned buffers over the C ABI
    model.py     `Model`: compile / fit / predict / save as utils/training.py uses them
    keras_api/   Adam, ModelCheckpoint, EarlyStopping, load_model, to_categorical
    utils/       mirror of the reference's utils/{deep_nn_models,training,preprocessing,performance_metrics}.py
    shims/       `keras` / `tensorflow` import names for running the reference's tune_*.py unchanged
    parallel.py  batch-sharded data-parallel trainer (NCCL) and the one-task-per-GPU sweep scheduler

Import it as `s2s_ismr_unet_b200` (alias package at the repo root)."""
