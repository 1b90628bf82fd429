import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from bench import synth_dataset
from s2s_ismr_unet_b200 import model as s2s_model
from s2s_ismr_unet_b200.keras_api.optimizers import Adam
from s2s_ismr_unet_b200.keras_api.callbacks import EarlyStopping
xt, yt, _ = synth_dataset(261 + 65, 64, 64, 1, seed=77)
xtr, ytr, xva, yva = xt[:261], yt[:261], xt[261:], yt[261:]
mc = s2s_model.Model((64, 64, 1), filters=2, n_blocks=3, ct_kernel=3, max_batch=32)
mc.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy", metrics=["accuracy"])
mc.fit(x=xtr, y=ytr, validation_data=(xva, yva), epochs=1, batch_size=16, shuffle=True, verbose=0)
for rep in range(3):
    t0 = time.perf_counter()
    mc.fit(x=xtr, y=ytr, validation_data=(xva, yva), epochs=5, batch_size=16, shuffle=True, verbose=0,
           callbacks=[EarlyStopping(monitor="val_loss", patience=10, restore_best_weights=True)])
    print("fit ms/epoch", 1e3 * (time.perf_counter() - t0) / 5)
mc.predict(xtr, verbose=0)
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(5):
        pg = mc.predict(xtr, verbose=0)
    print("predict samples/s", 5 * len(xtr) / (time.perf_counter() - t0))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
mc.fit(x=xtr, y=ytr, validation_data=(xva, yva), epochs=5, batch_size=16, shuffle=True, verbose=0)
for _ in range(5): mc.predict(xtr, verbose=0)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
