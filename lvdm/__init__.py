"""Import-path aliases: the reference's `lvdm.*` targets (configs/inference_512_v1.0.yaml, scripts/evaluation/*.py,
gradio_app.py) resolve to the B200-native implementation in `tooncrafter_b200`.  No arithmetic lives here."""
