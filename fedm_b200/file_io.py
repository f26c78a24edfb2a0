"""Input readers of the problem-definition API (host side, dolfin-free).

Same names, arguments, return values and error behaviour as the reference's ``fedm/file_io.py`` (readers at
:226-521, ``Files`` at :34-117) so that example scripts keep working; implemented independently with ``re`` and
numpy only.  Output writers (``file_output``, ``log`` ...) are not part of the hot path.
"""
import itertools
import re
from pathlib import Path

import numpy as np

from .utils import comma_separated, print_rank_0


class Files:
    """Input/output locations (``fedm.files``): ``file_input`` must exist, ``output_folder_path`` is created on demand,
    ``error_file`` / ``model_log`` are truncated the first time they are requested for an output folder."""

    def __init__(self):
        self._inp = Path.cwd() / "file_input"
        self._out = Path.cwd() / "output"
        self._fresh = {"relative error.log": True, "model.log": True}

    file_input = property(lambda self: self._inp)

    @file_input.setter
    def file_input(self, value):
        value = Path(value)
        if not value.is_dir():
            raise RuntimeError(f"fedm.files.file_input: '{value}' is not a directory")
        self._inp = value

    output_folder_path = property(lambda self: self._out)

    @output_folder_path.setter
    def output_folder_path(self, value):
        value = Path(value)
        if value.resolve() != self._out.resolve():
            self._fresh = {k: True for k in self._fresh}
        if not value.is_dir():
            value.mkdir()
        self._out = value

    def _log_path(self, name):
        path = self._out / name
        if self._fresh[name]:
            path.parent.mkdir(parents=True, exist_ok=True)
            path.write_text("")
            self._fresh[name] = False
        return path

    error_file = property(lambda self: self._log_path("relative error.log"))
    model_log = property(lambda self: self._log_path("model.log"))


files = Files()


def decomment(lines):
    """Yields every line stripped of its '#' comment; blank results are skipped."""
    for raw in lines:
        text = raw.partition("#")[0].strip()
        if text:
            yield text


def read_and_decomment(file_name):
    with open(file_name, "r") as fh:
        return list(decomment(fh))


def read_single_value(file_name, convert=lambda x: x):
    with open(file_name, "r", encoding="utf8") as fh:
        for text in decomment(fh):
            return convert(text)
    raise RuntimeError(f"fedm.read_single_value: No value found in file '{file_name}'")


def read_single_float(file_name, convert=None):
    return read_single_value(file_name, convert=float)


def read_single_string(file_name):
    return read_single_value(file_name, convert=str)


def read_two_columns(file_name):
    rows = [[float(tok) for tok in text.split()[:2]] for text in read_and_decomment(file_name)]
    table = np.asarray(rows, dtype=np.float64)
    return list(table[:, 0]), list(table[:, 1])


def flatten(input_list):
    return list(itertools.chain.from_iterable(input_list))


def flatten_float(input_list):
    return [float(x) for x in flatten(input_list)]


def read_speclist(file_path):
    """(number of species, names, property file names, file-name stems) from 'speclist.cfg'."""
    entries = [text.replace("file:", "").split() for text in read_and_decomment(Path(file_path) / "speclist.cfg")
               if "file:" in text]
    names = [e[0] for e in entries]
    prop_files = [e[1] for e in entries]
    stems = [e[1].split(".")[0] for e in entries]
    return len(names), names, prop_files, stems


def reaction_matrices(path, species):
    """Power, loss and gain matrices of 'reacscheme.cfg' by counting species names on each side of ' -> '."""
    reactions = [text.partition(" Type:")[0] for text in read_and_decomment(Path(path) / "reacscheme.cfg")]
    lhs = [r.partition(" -> ")[0].rstrip() for r in reactions]
    rhs = [r.partition(" -> ")[2].rstrip() for r in reactions]
    consumed = np.array([[side.count(s) for s in species] for side in lhs], dtype=int).reshape(len(reactions), len(species))
    produced = np.array([[side.count(s) for s in species] for side in rhs], dtype=int).reshape(len(reactions), len(species))
    net = consumed - produced
    return consumed, np.where(net > 0, net, 0), np.where(net < 0, -net, 0)


def rate_coefficient_file_names(path):
    found = flatten(re.findall(r"kfile: ([A-Za-z0-9_]+.[A-Za-z0-9_]+)", text)
                    for text in read_and_decomment(Path(path) / "reacscheme.cfg"))
    return [Path(path) / "rate_coefficients" / name for name in found]


def read_energy_loss(path):
    pattern = re.compile(r"Uin:\s?([+-]?\d+.\d+[eE]?[-+]?\d+|0|1.0)")
    values = flatten_float(pattern.findall(text) for text in read_and_decomment(Path(path) / "reacscheme.cfg"))
    print_rank_0(values)
    return values


def read_dependence(file_name):
    file_name = Path(file_name)
    if not file_name.is_file():
        raise FileNotFoundError(f"fedm.read_dependence: file '{file_name}' not found")
    with open(file_name, "r", encoding="utf8") as fh:
        for raw in fh:
            if "Dependence:" in raw:
                return raw.split()[2]
    raise RuntimeError(f"fedm.read_dependence: Did not find dependence in file '{file_name}'")


def read_dependences(file_names, zero_if_file_missing=False):
    out = []
    for name in file_names:
        try:
            out.append(read_dependence(name))
        except FileNotFoundError:
            if not zero_if_file_missing:
                raise
            out.append(0)
    return out


_RATE_KINDS = {"const": "float", "fun:Te,Tgas": "str", "fun:Tgas": "str", "Umean": "table", "E/N": "table", "ElecDist": "table"}


def read_rate_coefficients(rc_file_names, k_dependences):
    if len(rc_file_names) != len(k_dependences):
        raise ValueError("fedm.read_rate_coefficients: rc_file_names and k_dependences should be the same length.")
    for dep in k_dependences:
        if dep not in _RATE_KINDS:
            raise ValueError(f"fedm.read_rate_coefficients: The dependence '{dep}' is not recognised. "
                             f"Options are {comma_separated(list(_RATE_KINDS))}.")
    kxs, kys = [], []
    for dep, name in zip(k_dependences, rc_file_names):
        print_rank_0(name)
        kind = _RATE_KINDS[dep]
        if kind == "table":
            kx, ky = read_two_columns(name)
        elif kind == "float":
            kx, ky = 0.0, read_single_float(name)
        else:
            kx, ky = 0.0, read_single_string(name)
        kxs.append(kx)
        kys.append(ky)
    return kxs, kys


def read_transport_coefficients(particle_names, transport_type, model):
    folder = files.file_input / model / "transport_coefficients"
    if not folder.is_dir():
        raise FileNotFoundError(f"fedm.read_transport_coefficients: Transport coeff dir '{folder}' not found.")
    kinds = {"const": "float", "const.": "float", "fun:Te,Tgas": "str", "fun:E": "str",
             "Umean": "table", "E/N": "table", "Tgas": "table", "Te": "table"}
    mobility = transport_type == "mobility"
    if transport_type == "Diffusion":
        kinds["ESR"] = "esr"
    if mobility:
        kinds[0] = "missing"
    suffix = "_ND.dat" if transport_type == "Diffusion" else "_Nb.dat"
    paths = [folder / (name + suffix) for name in particle_names]
    deps = read_dependences(paths, zero_if_file_missing=mobility)
    for dep in deps:
        if dep not in kinds:
            raise ValueError(f"fedm.read_transport_coefficients: Dependence '{dep}' not recognised. For the transport type "
                             f"'{transport_type}', the possible options are {comma_separated(list(kinds))}.")
    kxs, kys = [], []
    for path, dep in zip(paths, deps):
        kind = kinds[dep]
        if kind == "missing":
            kxs.append(0)
            kys.append(0)
            continue
        print_rank_0(path)
        if kind == "table":
            kx, ky = read_two_columns(path)
        elif kind == "esr":
            kx, ky = 0.0, 0.0
        elif kind == "float":
            kx, ky = 0.0, read_single_float(path)
        else:
            kx, ky = 0.0, read_single_string(path)
        if dep == "fun:Te,Tgas":
            text = ky
            try:
                ky = eval(ky)
            except Exception as exc:
                raise RuntimeError(f"fedm.read_transport_coefficients: ky eval failed, '{text}'") from exc
        kxs.append(kx)
        kys.append(ky)
    return kxs, kys, deps


def read_particle_properties(file_names, model):
    folder = files.file_input / model / "species"
    mass_re = re.compile(r"Mass\s?=\s?([+-]?\d+.\d+[eE]?[-+]?\d+|0|1.0)")
    charge_re = re.compile(r"Z\s+?=\s+?([+-]?\d+)")
    masses, charges = [], []
    for name in file_names:
        path = folder / name
        if not path.is_file():
            raise RuntimeError(f"fedm.read_particle_properties: File '{path}' not found.")
        print_rank_0(path)
        got_mass = got_charge = False
        for text in read_and_decomment(path):
            print_rank_0(text)
            m, z = mass_re.findall(text), charge_re.findall(text)
            if m:
                masses.append(float(m[0]))
                got_mass = True
            if z:
                charges.append(float(z[0]))
                got_charge = True
        if not got_mass:
            raise RuntimeError(f"fedm.read_particle_properties: No mass found in file '{path}'.")
        if not got_charge:
            raise RuntimeError(f"fedm.read_particle_properties: No charge found in file '{path}'.")
    return masses, charges


def print_time_step(dt):
    print_rank_0("Time step is dt =", dt)


def print_time(t):
    print_rank_0("t =", t)
