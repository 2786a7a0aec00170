"""`python -m birdnet_stm32 <command>` -- only the hot-path command exists in this package."""

import sys


def main():
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help"):
        print("usage: python -m birdnet_stm32 evaluate [options]   (B200 engine; train/convert/deploy stay in the reference)")
        return 0 if len(sys.argv) >= 2 else 2
    cmd, rest = sys.argv[1], sys.argv[2:]
    if cmd == "evaluate":
        from birdnet_stm32.cli.evaluate import main as run

        run(rest)
        return 0
    if cmd == "export-blob":
        from birdnet_stm32.conversion.export_blob import export_blob_file

        if len(rest) < 2:
            print("usage: python -m birdnet_stm32 export-blob <model.tflite> <out.b200blob> [model_config.json]")
            return 2
        print(export_blob_file(rest[0], rest[2] if len(rest) > 2 else None, rest[1]))
        return 0
    print(f"unknown command '{cmd}' (this package implements: evaluate, export-blob)")
    return 2


if __name__ == "__main__":
    sys.exit(main())
