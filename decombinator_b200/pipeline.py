"""Entry point: ``decombinator {decombine|pipeline|collapse|translate}`` (reference pipeline.py:41-53)."""
import sys

from .decombine import decombinator
from .io import cli_args, write_out_intermediate


def _later_stage(name):
    print("decombinator_b200: the '%s' stage is not part of this build yet (the decombine hot path is); "
          "run it with the reference package on the .n12 written by `decombine`." % name)
    sys.exit(2)


def run(args=None, cli_args=None):
    """Decombine, write the .n12, then hand over to the later stages (reference pipeline.py:10-38)."""
    inputargs = cli_args if cli_args else args
    data = decombinator(inputargs)
    if not inputargs["dontsave"]:
        write_out_intermediate(data, inputargs, ".n12")
    print("Decombinator complete...")
    try:
        from .collapse import collapsinator
    except ImportError:
        _later_stage("collapse")
    data = collapsinator(data=data, inputargs=inputargs)
    if not inputargs["dontsave"]:
        write_out_intermediate(data, inputargs, ".freq")
    print("Collapsinator complete...")
    return data


def main():
    inputargs = cli_args()
    if inputargs["command"] == "decombine":
        data = decombinator(inputargs)
        write_out_intermediate(data, inputargs, ".n12")
    elif inputargs["command"] == "collapse":
        try:
            from .collapse import collapsinator
        except ImportError:
            _later_stage("collapse")
        data = collapsinator(inputargs=inputargs)
        write_out_intermediate(data, inputargs, ".freq")
    elif inputargs["command"] == "translate":
        _later_stage("translate")
    else:
        run(cli_args=inputargs)


if __name__ == "__main__":
    main()
