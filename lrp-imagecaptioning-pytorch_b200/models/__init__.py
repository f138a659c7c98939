"""Host-side mirrors of the reference's model-level LRP entry points (models/gridTDmodel.py,
models/aoamodel.py) and of the encoder definitions they wrap (models/vgg.py, models/resnet.py)."""
