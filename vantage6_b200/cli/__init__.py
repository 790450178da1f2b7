"""vnode / vserver command line interface (parity target: reference vantage6/cli/*)."""
from .._version import __version__, version_info  # noqa: F401
