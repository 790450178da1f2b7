"""Import location kept for code written against ``vantage6.cli.utils``; the guards themselves live with the
rest of the shared CLI machinery in :mod:`vantage6_b200.cli.instance`."""
from .instance import check_config_name_allowed, check_if_docker_deamon_is_running

__all__ = ["check_config_name_allowed", "check_if_docker_deamon_is_running"]
