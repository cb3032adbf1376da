"""`import pgdvs_b200.compat as pytorch3d` — the pytorch3d names the PGDVS renderers touch
(/root/reference/pgdvs/renderers/pgdvs_renderer_dyn.py:9-11, 405-410, 684-722;
st_geo_renderer.py:10-12, 37-42, 85-120), backed by the sm_100a kernels.  See INTEGRATION.md."""
from . import ops, renderer, structures, utils  # noqa: F401
