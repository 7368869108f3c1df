"""Minimal attribute-access config (stand-in for omegaconf.DictConfig, which is not installed offline) and the
reference's default model / diffusion configurations (configs/model/{cdm,cmdm}.yaml, configs/default.yaml:31-40)."""
import copy


class AttrDict(dict):
    """dict with recursive attribute access — the models only ever do `cfg.key` reads (cdm.py:418-472, cmdm.py:19-76)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v

    def __setattr__(self, k, v):
        self[k] = v


def cdm_model_cfg(num_points: int = 8192, input_feats: int = 6, use_scene_model: bool = False, use_color: bool = False) -> AttrDict:
    """configs/model/cdm.yaml with the H3D overrides of scripts/t2m_contact/*.sh (arch=Perceiver, no scene model);
    use_scene_model=True gives the HUMANISE / novel variant with the frozen PointTransformerSeg (cdm.yaml:17-25)."""
    return AttrDict(copy.deepcopy(dict(
        name="CDM", input_feats=input_feats, data_repr="contact_cont_joints", time_emb_dim=128,
        text_model=dict(version="ViT-B/32", max_length=32),
        scene_model=dict(name="PointTransformerSeg", use_scene_model=use_scene_model, use_color=use_color, use_openscene=False,
                         num_points=num_points, point_feat_dim=32, pretrained_weight=None, freeze=True),
        arch="Perceiver",
        arch_perceiver=dict(last_dim=256, point_pos_emb=True, encoder_q_input_channels=512, encoder_kv_input_channels=256,
                            encoder_num_heads=8, encoder_widening_factor=1, encoder_dropout=0.1, encoder_residual_dropout=0.0,
                            encoder_self_attn_num_layers=2, decoder_q_input_channels=256, decoder_kv_input_channels=512,
                            decoder_num_heads=8, decoder_widening_factor=1, decoder_dropout=0.1, decoder_residual_dropout=0.0),
    )))


def cmdm_model_cfg(num_points: int = 8192, input_feats: int = 263) -> AttrDict:
    """configs/model/cmdm.yaml with the H3D overrides of scripts/t2m_contact_motion/*.sh."""
    return AttrDict(copy.deepcopy(dict(
        name="CMDM", input_feats=input_feats, data_repr="h3d", time_emb_dim=512,
        contact_model=dict(contact_type="contact_cont_joints", contact_joints=[0, 10, 11, 12, 20, 21], planes=[32, 64, 128, 256],
                           num_points=num_points, blocks=[2, 2, 2, 2]),
        text_model=dict(version="ViT-B/32", max_length=32),
        arch="trans_enc", latent_dim=512, mask_motion=True, num_layers=[1, 1, 1, 1, 1], num_heads=8, dropout=0.1,
        dim_feedforward=1024,
    )))


def diffusion_cfg(steps: int = 1000, timestep_respacing: str = "") -> AttrDict:
    """configs/default.yaml:31-40."""
    return AttrDict(dict(predict_xstart=True, steps=steps, noise_schedule="cosine", timestep_respacing=timestep_respacing,
                         rescale_timesteps=False, loss_type="MSE", learn_sigma=False, sigma_small=True))


def full_cfg(model: AttrDict, steps: int = 1000, timestep_respacing: str = "") -> AttrDict:
    return AttrDict(dict(model=model, diffusion=diffusion_cfg(steps, timestep_respacing)))
