"""Minimal colorama replacement (colorama is not installed): ANSI colours that are emitted
only when the stream is a TTY (``click.echo`` strips them otherwise, which is what the
reference's golden-output tests rely on: reference tests/test_node_cli.py:80-87)."""


class _Fore:
    RED = "\033[31m"
    GREEN = "\033[32m"
    YELLOW = "\033[33m"
    BLUE = "\033[34m"
    CYAN = "\033[36m"
    RESET = "\033[39m"


class _Style:
    BRIGHT = "\033[1m"
    RESET_ALL = "\033[0m"


Fore = _Fore()
Style = _Style()
