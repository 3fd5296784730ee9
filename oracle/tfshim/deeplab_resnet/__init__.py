"""Stand-in for external/tensorflow-deeplab-resnet's package -- TEST INFRASTRUCTURE ONLY (see ../tensorflow/__init__.py).
The backbone is out of scope (BASELINE.json north_star: "fed as precomputed synthetic tensors"); the reference reads three
taps from it (CMPC_model.py:73-76), which ``model.DeepLabResNetModel`` hands out from the tensors ``oracle/ref_runner.py`` fed."""
