"""TEST INFRASTRUCTURE ONLY -- numpy statement of the forcing-preparation kernel (noahmp_b200/csrc/forcing_prep.cuh).

PARITY UNPINNED: the reference has the namelist fields (offline/noahmp_config.py:81-82,98) but no interpolation or
solar-geometry code at all, so this is this repo's own specification, stated twice (numpy fp32 here, CUDA there)
and required to agree bit for bit."""
import numpy as np

f32 = np.float32


def prepare_forcing(rec0, rec1, weight, geo, scal4):
    """rec0, rec1 [8][ncol]; geo [4][ncol] = sin(lat) cos(lat) sin(lon) cos(lon); scal4 = sd, cd, ch0, sh0."""
    w = f32(weight)
    sd, cd, ch0, sh0 = (f32(x) for x in scal4)
    out = np.empty((9, rec0.shape[1]), f32)
    out[:8] = rec0 + w * (rec1 - rec0)
    sinlat, coslat, sinlon, coslon = geo
    ch = ch0 * coslon - sh0 * sinlon
    out[8] = np.minimum(np.maximum(sinlat * sd + (coslat * cd) * ch, f32(-1.0)), f32(1.0))
    return out
