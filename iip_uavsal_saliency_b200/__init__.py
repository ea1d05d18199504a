"""uavsal-b200: the UAVSal per-frame video-saliency inference path and its CC/NSS/KLD/SIM metrics as hand-written
sm_100a kernels behind the reference's PyTorch-facing surface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
