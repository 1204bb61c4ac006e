"""Stand-in for the reference's ColorFusionResidualNet with the same layer shapes and attribute names
(color_aggregation_network.py:6-49,71-100), for tools that must run without the staged reference files."""
import torch


class ConvDecoderAE(torch.nn.Module):      # same layer shapes as color_aggregation_network.py:6-49 (38 channels)
    def __init__(self, h=38):
        super().__init__()
        c = lambda i, o, k=3: torch.nn.Sequential(torch.nn.Conv2d(i, o, k, padding=k // 2), torch.nn.ReLU())
        self.enc1, self.enc2, self.enc3 = c(h, h), c(h, h // 2), c(h // 2, h // 4)
        self.up2_conv, self.up1_conv = c(h // 4, h // 2), c(h // 2, h)
        self.dec2, self.dec1 = c(h, h // 2), c(2 * h, h)
        self.fuse_input = c(2 * h, h, 1)
        self.final = torch.nn.Conv2d(h, 3, 1)


class Net(torch.nn.Module):
    def __init__(self, mode):
        super().__init__()
        self.per_view_feat_dim, self.feat_aggregate_mode = 32, mode
        self.per_view_mlp = torch.nn.Sequential(torch.nn.Linear(7, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU())
        self.conv_decoder = ConvDecoderAE()


class Opts:
    enable_exposure_correction = False
    nb_visible_src_frames = 3
    residual_resolution_scale = 1.0


