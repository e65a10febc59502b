"""Names of the reference's utils/visualization.py (plotly / trimesh viewers) so that the notebooks' import cells work.
Viewers are outside this package's scope: the functions warn and return None."""
import warnings


def plot_voxel(*args, **kwargs):
    """visualization.py: interactive plotly scatter of a voxel grid -- not reproduced."""
    warnings.warn("plot_voxel: viewers are outside this package's scope; nothing is drawn "
                  "(utils.voxel_utils.voxel_grid_to_points exports the points)", stacklevel=2)


def visualize_mesh_plotly(*args, **kwargs):
    """visualization.py: plotly mesh viewer -- not reproduced."""
    warnings.warn("visualize_mesh_plotly: viewers are outside this package's scope; nothing is drawn", stacklevel=2)
