"""Entry point: ``decombinator {decombine|pipeline|collapse|translate}`` (reference pipeline.py:10-53).

Under ``torchrun`` (WORLD_SIZE > 1) the same commands run one process per GPU: reads are sharded over the ranks for
``decombine``, rows are repartitioned by barcode hash for ``collapse`` (parallel.py); rank 0 writes the files."""
import os
import sys

from .decombine import decombinator
from .io import cli_args, write_out_intermediate


def _translate_out_of_scope():
    print("decombinator_b200: the 'translate' stage is outside this build (DESIGN.md section 7: it runs on the collapsed, "
          "tiny output); run the reference's `decombinator translate` on the .freq file written here.")
    sys.exit(2)


def _world():
    return int(os.environ.get("WORLD_SIZE", "1"))


def run(args=None, cli_args=None):
    """Decombine -> .n12 -> collapse -> .freq, stages chained in memory (reference pipeline.py:10-38)."""
    inputargs = cli_args if cli_args else args
    from .collapse import collapsinator
    if _world() > 1:
        from . import parallel
        rank, _ = parallel.init_from_env()
        mine, first = parallel.decombinator_shard(inputargs)
        data = parallel.gather_rows(mine)
        if rank == 0 and not inputargs["dontsave"]:
            write_out_intermediate(data, inputargs, ".n12")
        print("Decombinator complete...")
        data = parallel.collapsinator_sharded(inputargs, data=mine, first_index=first)
        if rank == 0 and not inputargs["dontsave"]:
            write_out_intermediate(data, inputargs, ".freq")
        print("Collapsinator complete...")
        return data
    data = decombinator(inputargs)
    if not inputargs["dontsave"]:
        write_out_intermediate(data, inputargs, ".n12")
    print("Decombinator complete...")
    data = collapsinator(data=data, inputargs=inputargs)
    if not inputargs["dontsave"]:
        write_out_intermediate(data, inputargs, ".freq")
    print("Collapsinator complete...")
    return data


def main():
    inputargs = cli_args()
    multi = _world() > 1
    if multi:
        from . import parallel
        rank, _ = parallel.init_from_env()
    else:
        rank = 0
    if inputargs["command"] == "decombine":
        if multi:
            data = parallel.decombinator_sharded(inputargs)
        else:
            inputargs["rows_as_text"] = True      # nothing reads the rows but the .n12 writer
            data = decombinator(inputargs)
        if rank == 0:
            write_out_intermediate(data, inputargs, ".n12")
    elif inputargs["command"] == "collapse":
        from .collapse import collapsinator
        data = parallel.collapsinator_sharded(inputargs) if multi else collapsinator(inputargs=inputargs)
        if rank == 0:
            write_out_intermediate(data, inputargs, ".freq")
    elif inputargs["command"] == "translate":
        _translate_out_of_scope()
    else:
        run(cli_args=inputargs)


if __name__ == "__main__":
    main()
