"""Host-side writers of src/scripts/voxel_grid.py that the data-generation scripts call after the kernels
(raytracing.py:166 txt_to_mesh): the cube-per-voxel ASCII PLY used for eyeballing
grids. Pure file formatting, byte-compatible with the reference (including its indented header lines)."""
import numpy as np

from .config import vox_dim

_CUBE_VERTS = np.array([[-1.0, -1.0, 1.0], [1.0, -1.0, 1.0], [1.0, 1.0, 1.0], [-1.0, 1.0, 1.0],
                        [-1.0, -1.0, -1.0], [1.0, -1.0, -1.0], [1.0, 1.0, -1.0], [-1.0, 1.0, -1.0]])
_CUBE_FACES = np.array([[0, 1, 2], [2, 3, 0], [1, 5, 6], [6, 2, 1], [7, 6, 5], [5, 4, 7],
                        [4, 0, 3], [3, 7, 4], [4, 5, 1], [1, 0, 4], [3, 2, 6], [6, 7, 3]])


def write_ply(filename, verts, faces):
    """voxel_grid.py:263-283. verts: rows x y z r g b; faces: rows of 3 vertex indices."""
    with open(filename, "w") as f:
        f.write('''ply
    format ascii 1.0
    element vertex %d
    property float x
    property float y
    property float z
    property uchar red
    property uchar green
    property uchar blue
    element face %d
    property list uchar int vertex_index
    end_header
    ''' % (len(verts), len(faces)))
        f.write("".join('%f %f %f %d %d %d\n' % tuple(v) for v in verts))
        f.write("".join('3 %d %d %d\n' % tuple(t) for t in faces))


def grid_to_mesh(grid_lst, ply_file, grid_size=None):
    """voxel_grid.py:223-260: rows (i, j, k, r, g, b) -> one 0.9-voxel cube each."""
    rows = np.asarray(grid_lst).reshape(-1, 6)
    grid_size = 1 if grid_size is None else grid_size
    min_bound = np.array([-grid_size / 2, -grid_size / 2, -grid_size / 2])
    voxel_scale = grid_size / vox_dim
    verts, faces = [], []
    for n, data in enumerate(rows):
        ijk = data[:3].astype(int)
        col = data[3:6].astype(int)
        v = (_CUBE_VERTS * 0.45 + ijk).astype(float) * voxel_scale + min_bound
        verts += [list(p) + list(col) for p in v]
        faces += [list(8 * n + fc) for fc in _CUBE_FACES]
    write_ply(ply_file, verts, faces)
    return verts, faces


def txt_to_mesh(txt_file, ply_file, grid_size=None):
    """voxel_grid.py:180-220."""
    grid_to_mesh(np.loadtxt(txt_file, dtype=int, ndmin=2), ply_file, grid_size)
