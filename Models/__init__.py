"""Drop-in `Models` package: the import paths the reference's driver scripts use
(`INFERENCE.py:71`, `TRAIN_FINAL.py:14`, `INFERENCE_TIMER.py:177`)."""
