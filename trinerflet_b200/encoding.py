"""get_encoder with the reference's signature (reconstruction/encoding.py:45-96) for the two encodings on the
hot path: 'triplane_wavelet' (:75-93) and 'sphere_harmonics' (:60-62)."""
from .shencoder import SHEncoder
from .triplane_encoder import TriPlaneVolume


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16,
                log2_hashmap_size=19, desired_resolution=2048, align_corners=False, bound=1, **kwargs):
    if encoding == 'None':
        return (lambda x, **kw: x), input_dim
    if encoding == 'sphere_harmonics':
        encoder = SHEncoder(input_dim=input_dim, degree=degree)
    elif encoding == 'triplane_wavelet':
        encoder = TriPlaneVolume(
            number_of_features=kwargs['triplane_channels'],
            plane_resolution=kwargs['triplane_resolution'],
            init_sigma=0.1,
            lbound=bound,
            viewdir_plane_resolution=-1,
            apply_activation_on_features=False,
            inner_multi_res_scale=kwargs['triplane_wavelet_levels'],
            inner_multi_res_scale_current=1,
            learn_rotation_axis=kwargs.get('learn_rotation_axis', False),
            dropout=kwargs.get('dropout', 0),
            wavelet_type=kwargs.get('wavelet_type', 'bior6.8'),
            lbound_auto_scale=kwargs.get('lbound_auto_scale', False),
            upscale_ratio_bound=kwargs.get('upscale_ratio_bound', -1),
            upscale_levels=kwargs.get('upscale_levels', 2),
            wavelet_base_resolution=kwargs.get('wavelet_base_resolution', 0),
        )
    else:
        raise NotImplementedError(
            f"encoding '{encoding}' is outside the trinerflet_b200 hot path (SURVEY.md section 8: hashgrid/frequency/"
            "tiledgrid are never selected with --triplane_wavelet)")
    return encoder, encoder.output_dim
