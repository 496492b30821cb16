"""Birds_Eye_View_Loss variant of the reference's module surface (same kernels, normalised
[0,1) BEV coordinates, ``y = 1 - y'``, float32 beta, 2-tuple ERFNet output).  Put this
directory first on sys.path to serve the imports of ``Birds_Eye_View_Loss/main.py``."""
