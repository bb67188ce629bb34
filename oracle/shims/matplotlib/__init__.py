"""Stand-in for matplotlib (oracle only): the scenarios' visualize.py modules only need
`plt.cm.get_cmap` at construction time and `matplotlib.patches` to be importable; figures are
off (show_figure_frequency = -1) on the path under test."""
from . import patches  # noqa: F401
from . import pyplot  # noqa: F401
