"""``python -m paddle_edl.protos.run_codegen`` -- CLI parity with the reference's protoc driver
(python/edl/protos/run_codegen.py, generate.sh).  Nothing is generated here: the message classes are built at
import time from ``schema.py``; this command only refreshes the human-readable ``*.proto`` renderings."""
import os

from . import schema


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name in schema.PROTO_FILES:
        path = os.path.join(here, os.path.basename(name))
        with open(path, "w") as f:
            f.write(schema.render_proto(name))
        print("wrote", path)


if __name__ == "__main__":
    main()
