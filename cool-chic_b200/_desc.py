"""ctypes image of ``CcdCoolChicDesc`` (include/ccdec.h) and its construction from a parsed
``CoolChicHeader``.  Everything the device side needs to know about one Cool-chic's
architecture: grid sizes, IFCE inputs, ARM / upsampling / synthesis shapes, q-steps and
exp-Golomb orders.  Values follow the reference's header (header.py:243-377) and
``CoolChicEncoderParameter.__post_init__`` (core/coolchic.py:149-225)."""
import ctypes
import math

from .bitstream.header import FINAL_UPSAMPLING, NN_KINDS, NN_MODULES, CoolChicHeader

CCD_MAX_GRIDS = 32
CCD_MAX_SYN = 8


class CcdCoolChicDesc(ctypes.Structure):
    _fields_ = [
        ("img_h", ctypes.c_int32),
        ("img_w", ctypes.c_int32),
        ("n_grids", ctypes.c_int32),
        ("grid_h", ctypes.c_int32 * CCD_MAX_GRIDS),
        ("grid_w", ctypes.c_int32 * CCD_MAX_GRIDS),
        ("grid_is_hyper", ctypes.c_int32 * CCD_MAX_GRIDS),
        ("grid_ifce_in", ctypes.c_int32 * CCD_MAX_GRIDS),
        ("latent_res_lo", ctypes.c_int32),
        ("latent_res_hi", ctypes.c_int32),
        ("n_ctx", ctypes.c_int32),
        ("n_ifce_out", ctypes.c_int32),
        ("arm_hidden", ctypes.c_int32),
        ("arm_stab", ctypes.c_int32),
        ("ups_k", ctypes.c_int32),
        ("ups_pre_k", ctypes.c_int32),
        ("n_ups", ctypes.c_int32),
        ("n_syn_layers", ctypes.c_int32),
        ("syn_out", ctypes.c_int32 * CCD_MAX_SYN),
        ("syn_k", ctypes.c_int32 * CCD_MAX_SYN),
        ("syn_res", ctypes.c_int32 * CCD_MAX_SYN),
        ("syn_relu", ctypes.c_int32 * CCD_MAX_SYN),
        ("syn_stab", ctypes.c_int32),
        ("syn_in", ctypes.c_int32),
        ("common_randomness", ctypes.c_int32),
        ("final_ups", ctypes.c_int32),
        ("qshift", ctypes.c_int32 * 8),
        ("expgol", ctypes.c_int32 * 8),
        ("nn_n_bit_pad", ctypes.c_int32),
        ("flag_ifce", ctypes.c_int32),
    ]

    @property
    def n_out_channels(self) -> int:
        return int(self.syn_out[self.n_syn_layers - 1])

    def grid_sizes(self):
        return [(int(self.grid_h[i]), int(self.grid_w[i])) for i in range(self.n_grids)]

    def n_symbols(self) -> int:
        return sum(h * w for h, w in self.grid_sizes())


def desc_from_header(header: CoolChicHeader) -> CcdCoolChicDesc:
    p = header.get_coolchic_parameter()
    if p.n_latent_grids != header.get_value("n_latent_grids"):
        raise ValueError(
            f"Header announces n_latent_grids={header.get_value('n_latent_grids')} but its "
            f"resolutions give {p.n_latent_grids}."
        )
    if not 1 <= p.n_latent_grids <= CCD_MAX_GRIDS:
        raise ValueError(f"n_latent_grids={p.n_latent_grids} out of range")
    d = CcdCoolChicDesc()
    d.img_h, d.img_w = p.img_size
    d.n_grids = p.n_latent_grids
    for i, (size, hyp, nin) in enumerate(zip(p.size_per_latent, p.flag_is_hyperlatent, p.input_features_ifce)):
        d.grid_h[i], d.grid_w[i] = size[-2], size[-1]
        d.grid_is_hyper[i] = int(hyp)
        d.grid_ifce_in[i] = nin
    d.latent_res_lo, d.latent_res_hi = p.latent_resolution
    d.n_ctx = p.spatial_context_arm
    d.flag_ifce = int(p.flag_ifce)
    d.n_ifce_out = p.output_feature_ifce
    d.arm_hidden = p.n_hidden_layers_arm
    d.arm_stab = int(p.linear_stabiliser_arm)
    d.ups_k = p.ups_k_size
    d.ups_pre_k = p.ups_preconcat_k_size
    d.n_ups = p.latent_resolution[1]  # instantiate_ups_from_cc_param, core/coolchic.py:1081-1089
    layers = p.parsed_synthesis_layers()
    d.n_syn_layers = len(layers)
    for i, (o, k, res, relu) in enumerate(layers):
        d.syn_out[i], d.syn_k[i], d.syn_res[i], d.syn_relu[i] = o, k, int(res), int(relu)
    d.syn_stab = int(p.linear_stabiliser_synth)
    d.syn_in = p.input_feature_synthesis
    d.common_randomness = int(p.flag_common_randomness)
    d.final_ups = FINAL_UPSAMPLING.index(p.final_upsampling_type)
    q = header.get_value("nn_q_step")
    e = header.get_value("nn_expgol_cnt")
    j = 0
    for m in NN_MODULES:
        for k in NN_KINDS:
            d.qshift[j] = int(round(math.log2(q.get_value(m, k))))
            d.expgol[j] = int(e.get_value(m, k))
            j += 1
    d.nn_n_bit_pad = header.get_value("nn_n_bit_pad")
    return d
