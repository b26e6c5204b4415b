"""Constants of the reference's src/scripts/config.py:1-12, plus the two names
raytracing.py:31-32 reads but the reference forgot to define (render_img_*)."""
depth_factor = 100

# rendering images
img_width = 512
img_height = 512
render_img_width = 512
render_img_height = 512
znear = 0.4
zfar = 4.0
focal = 525.0
cam_depth = 1.4
color_bg = [1., 1., 1.]

vox_dim = 32

# raytracing.py:75 clamps pixel coordinates with the literal 511
index_clamp_max = 511
