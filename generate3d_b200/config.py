"""Configuration surface of the hot path.

The reference keeps these as bare module constants (src/scripts/config.py:1-12) and additionally reads two
names it never defines (`render_img_width` / `render_img_height`, raytracing.py:31-32). The drop-in modules
here look the values up as attributes of this module, exactly like the reference's `config.<name>`, so a
caller can still do `config.focal = 262.5` before a call. Values and their provenance:
"""

_DEFAULTS = {
    # voxel grids
    "vox_dim": 32,                 # config.py:12 -- every grid of the pipeline is vox_dim^3
    # depth / colour renderings (renderer.py writes them, raytracing.py reads them)
    "depth_factor": 100,           # config.py:1 -- uint8 depth = metres * depth_factor (renderer.py:34-40)
    "img_width": 512,              # config.py:4
    "img_height": 512,             # config.py:5
    "render_img_width": 512,       # read by raytracing.py:31, never defined upstream; same as img_width
    "render_img_height": 512,      # read by raytracing.py:32, never defined upstream; same as img_height
    "focal": 525.0,                # config.py:8
    "znear": 0.4,                  # config.py:6
    "zfar": 4.0,                   # config.py:7
    "cam_depth": 1.4,              # config.py:9 -- load_pose overwrites pose[2,3] with -cam_depth
    "color_bg": [1.0, 1.0, 1.0],   # config.py:10
    # raytracing.py:75 clamps projected pixel coordinates with the literal 511, whatever the image size
    "index_clamp_max": 511,
}

globals().update(_DEFAULTS)


def reset():
    """Restore every value above (tests that tweak focal / image size call this)."""
    globals().update(_DEFAULTS)
