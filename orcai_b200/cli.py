"""``orcai`` command line: the prediction-path subcommands of the reference CLI
(``src/orcAI/cli.py:93-237, 359-416``) with identical options, defaults and file outputs.

The training / data-preparation subcommands of the reference (create-label-arrays,
create-snippet-table, create-tvt-*, hpsearch, train, test, init, create-recording-table) are out of
scope for this package and are not registered.
"""

from __future__ import annotations

from importlib.resources import files
from pathlib import Path

import click

from orcai_b200 import __version__
from orcai_b200.auxiliary import Messenger

ClickDirPathR = click.Path(exists=True, file_okay=False, readable=True, resolve_path=True, path_type=Path)
ClickDirPathW = click.Path(exists=True, file_okay=False, writable=True, resolve_path=True, path_type=Path)
ClickDirPathWcreate = click.Path(exists=False, file_okay=False, writable=True, resolve_path=True, path_type=Path)
ClickFilePathR = click.Path(exists=True, dir_okay=False, readable=True, resolve_path=True, path_type=Path)

INCLUDED_MODELS = [f.name for f in files("orcai_b200.models").iterdir() if f.is_dir() and not f.name.startswith(("_", "."))]

EPILOG = "For further information visit: https://github.com/ethz-tb/orcAI"


@click.group(
    help="Command line interface for orcai_b200 - the B200-native orcAI prediction path "
    "(detects acoustic signals in spectrograms generated from audio recordings).",
    epilog="For further information see the help pages of the individual subcommands (e.g. orcai predict --help).",
)
@click.version_option(__version__, prog_name="orcai_b200")
def cli():
    pass


@cli.command(
    name="predict",
    help="Predicts call annotations from RECORDING_PATH. This can either be a path to a wav file or a recording table (created with create-recording-table) as .csv.",
    short_help="Predicts call annotations.",
    no_args_is_help=True,
    epilog=EPILOG,
)
@click.argument("recording_path", type=ClickFilePathR)
@click.option("--channel", "-c", type=int, default=1, show_default=True, help="Channel to use for prediction if running predicitons for a single file.")
@click.option(
    "--model",
    "-m",
    type=click.Choice(INCLUDED_MODELS, case_sensitive=False),
    default="orcai-v1",
    show_default=True,
    help="Builtin model to use for prediction. Overriden if model_dir is given.",
)
@click.option("--model_dir", "-md", "model_dir", type=ClickDirPathR, default=None, show_default="use builtin model", help="Path to a model directory.")
@click.option(
    "--output_path",
    "-o",
    default="default",
    show_default="default",
    help="Path to the output file/folder or 'default' to save in the same directory as the wav file. None to not save predictions to disk.",
)
@click.option("--overwrite", "-ow", is_flag=True, help="Overwrite existing predictions.")
@click.option("--save_probabilities", "-sp", is_flag=True, help="If True the prediction probabilities are saved to a file.")
@click.option(
    "--base_dir_recording",
    "-bdr",
    type=ClickDirPathW,
    default=None,
    show_default="None",
    help="Alternative base directory containing the recordings (possibly in subdirectories). If None the base directory is taken from the recording_table.",
)
@click.option(
    "--call_duration_limits",
    "-cdl",
    type=ClickFilePathR,
    default=None,
    show_default="None",
    help="Path to a JSON file containing call duration limits. None for no filtering based on call duration.",
)
@click.option("--label_suffix", "-ls", default="*", show_default=True, help="Suffix to add to the label names.")
@click.option("--verbosity", "-v", type=click.IntRange(0, 3), default=2, show_default=True, help="Verbosity level. 0: Errors only, 1: Warnings, 2: Info, 3: Debug")
def cli_predict(**kwargs):
    kwargs["msgr"] = Messenger(verbosity=kwargs["verbosity"], title="Predicting calls")
    from orcai_b200.predict import predict

    if kwargs["model_dir"] is None:
        kwargs["model_dir"] = files("orcai_b200.models").joinpath(kwargs["model"])
    del kwargs["model"]
    predict(**kwargs)


@cli.command(
    name="filter-predictions",
    help="Filters predictions in the predictions file at PREDICTION_FILE_PATH.",
    short_help="Filters predictions.",
    no_args_is_help=True,
    epilog=EPILOG,
)
@click.argument("predicted_labels", type=ClickFilePathR)
@click.option(
    "--call_duration_limits",
    "-cdl",
    type=ClickFilePathR,
    default=files("orcai_b200.defaults").joinpath("default_call_duration_limits.json"),
    show_default="default_call_duration_limits.json",
    help="Path to a JSON file containing call duration limits.",
)
@click.option("--output_file", "-o", default="default", show_default="default", help="Path to the output file or 'default' to save in the same directory as the prediction file.")
@click.option("--overwrite", "-ow", is_flag=True, help="Overwrite existing predictions.")
@click.option("--label_suffix", "-ls", default="*", show_default="*", help="Suffix to add to the label names.")
@click.option("--verbosity", "-v", type=click.IntRange(0, 3), default=2, show_default=True, help="Verbosity level. 0: Errors only, 1: Warnings, 2: Info, 3: Debug")
def cli_filter_predictions(**kwargs):
    kwargs["msgr"] = Messenger(verbosity=kwargs["verbosity"], title="Filtering predictions")
    from orcai_b200.predict import filter_predictions_file

    filter_predictions_file(**kwargs)


@cli.command(
    name="create-spectrograms",
    help="Creates spectrograms for all files in recording table at RECORDING_TABLE_PATH and writes them to OUTPUT_DIR.",
    short_help="Creates spectrograms.",
    no_args_is_help=True,
    epilog=EPILOG,
)
@click.argument("recording_table_path", type=ClickFilePathR)
@click.argument("output_dir", type=ClickDirPathWcreate)
@click.option(
    "--base_dir_recording",
    "-bdr",
    type=ClickDirPathR,
    default=None,
    show_default="None",
    help="Base directory for the wav files. If None the base_dir_recording is taken from the recording_table.",
)
@click.option(
    "--orcai_parameter",
    "-p",
    type=ClickFilePathR,
    default=files("orcai_b200.defaults").joinpath("default_orcai_parameter.json"),
    show_default="default_orcai_parameter.json",
    help="Path to the OrcAI parameter file.",
)
@click.option("--include_not_annotated", "-en", is_flag=True, help="Include recordings without annotations.")
@click.option("--include_no_possible_annotations", "-enp", is_flag=True, help="Include recordings without possible annotations.")
@click.option("--overwrite", "-ow", is_flag=True, help="Recreate existing spectrograms.")
@click.option("--verbosity", "-v", type=click.IntRange(0, 3), default=2, show_default=True, help="Verbosity level. 0: Errors only, 1: Warnings, 2: Info, 3: Debug")
def cli_create_spectrograms(**kwargs):
    kwargs["msgr"] = Messenger(verbosity=kwargs["verbosity"], title="Creating spectrograms")
    from orcai_b200.spectrogram import create_spectrograms

    create_spectrograms(**kwargs)


if __name__ == "__main__":
    cli()
