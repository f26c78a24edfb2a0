"""Input readers of the problem-definition API (host side, dolfin-free).

Same names, arguments, return values and error behaviour as the reference's ``fedm/file_io.py`` (readers at
:226-521, ``Files`` at :34-117) so that example scripts keep working; implemented independently with ``re`` and
numpy only.  Output writers (``file_output``, ``log`` ...) are not part of the hot path.
"""
import itertools
import re
from pathlib import Path

import numpy as np

from .utils import comma_separated, mesh_info, print_rank_0


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


def no_convert(x):
    """Identity, for readers that optionally convert what they read (fedm/file_io.py:123-129)."""
    return x


def read_single_value(file_name, convert=no_convert):
    with open(file_name, "r", encoding="utf8") as fh:
        for text in decomment(fh):
            return convert(text)
    raise RuntimeError(f"fedm.read_single_value: No value found in file '{file_name}'")


def read_single_float(file_name, convert=no_convert):
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


# ---- output (thin: PVD/VTU through the shim's File, checkpoints through its XDMFFile stand-in) -----------------------------
def truncate_file(path):
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text("")


def output_files(file_type, type_of_output, output_file_names):
    """One PVD or XDMF output object per name under ``files.output_folder_path/<type_of_output>/<name>/``."""
    from . import dolfin as df
    makers = {"pvd": df.File, "xdmf": df.XDMFFile}
    if file_type not in makers:
        raise ValueError(f"fedm.output_files: file type '{file_type}' is not valid. Options are 'pvd' or 'xdmf'.")
    folder = files.output_folder_path / type_of_output
    out = []
    for name in output_file_names:
        handle = makers[file_type](str(folder / name / f"{name}.{file_type}"))
        if file_type == "xdmf":
            handle.parameters["flush_output"] = True
        out.append(handle)
    return out


def file_output(t, t_old, t_out, step, t_out_list, step_list, file_type, output_file_list, particle_name, u_old, u_old1, unit="s"):
    """Writes every output instant t_out <= t, linearly interpolated between the two last time levels, and advances the
    output clock through the (t_out_list, step_list) schedule.  Returns (next t_out, current step)."""
    from . import dolfin as df
    scales = {"ns": 1e9, "us": 1e6, "ms": 1e3, "s": 1.0}
    if unit not in scales:
        raise ValueError(f"fedm.file_output: unit '{unit}' not valid. Options are {comma_separated(list(scales))}.'")
    scale = scales[unit]
    index = len(t_out_list) - 1 if t > max(t_out_list) else next(i for i, val in enumerate(t_out_list) if val > t)
    while t_out <= t:
        for i, handle in enumerate(output_file_list):
            snap = df.Function(u_old1[i].function_space())
            snap.vector()[:] = u_old1[i].vector()[:] + (t_out - t_old) * (u_old[i].vector()[:] - u_old1[i].vector()[:]) / (t - t_old)
            if file_type[i] == "pvd":
                snap.rename(particle_name[i], str(i))
                handle << (snap, t_out * scale)
            elif file_type[i] == "xdmf":
                handle.write_checkpoint(snap, particle_name[i], t_out * scale, df.XDMFFile.Encoding.HDF5, True)
            else:
                raise ValueError(f"fedm.file_output: file type '{file_type}' not recognised. Options are 'pvd' and 'xdmf'.")
        if 0.999 * t_out_list[index - 1] <= t_out < 0.999 * t_out_list[index]:
            step = step_list[index - 1]
        elif t_out >= 0.999 * t_out_list[index]:
            step = step_list[index]
        t_out += step
    return t_out, step


def mesh_statistics(mesh):
    from . import dolfin as df
    folder = files.output_folder_path / "mesh"
    df.File(str(folder / "mesh.pvd")).write(mesh)
    text = mesh_info(mesh)
    print_rank_0(text.rstrip())
    if df.MPI.rank(df.MPI.comm_world) == 0:
        (folder / "mesh info.txt").write_text(text)


def numpy_2d_array_to_str(x):
    body = str(np.asarray(x)).replace("[", "").replace("]", "")
    return "\n".join(line.strip() for line in body.split("\n"))


def log(log_type, log_file_name, *args):
    """Appends a block to model.log: 'properties', 'conditions', 'matrices', 'initial time', 'time', 'mesh'."""
    from . import dolfin as df
    if df.MPI.rank(df.MPI.comm_world) != 0:
        return
    kinds = ["properties", "conditions", "matrices", "initial time", "time", "mesh"]
    if log_type not in kinds:
        raise ValueError(f"fedm.log: log_type '{log_type}' not valid. Options are {comma_separated(kinds)}.")
    with open(log_file_name, "a") as fh:
        if log_type == "properties":
            gas, model, names, masses, charges = args
            fh.write(f"Gas:\t{gas}\n\nmodel:\t{model}\n\nParticle names: \n{names}\n\nMass:\n{masses}\n\nCharge:\n{charges}\n\n")
        elif log_type == "conditions":
            dt_var, U_w, p0, gap, N0, Tgas = args
            fh.write(f"Simulation conditions:\ndt = {dt_var} s\tU_w = {U_w} V\tp_0 = {p0} Torr\td = {gap} m\tN_0 = {N0} m^-3\tT_gas = {Tgas} K\n\n")
        elif log_type == "matrices":
            gain, loss, power = args
            fh.write("gain matrix:\n" + numpy_2d_array_to_str(gain) + "\n\nloss matrix:\n" + numpy_2d_array_to_str(loss) +
                     "\n\npower matrix:\n" + numpy_2d_array_to_str(power) + "\n\n")
        elif log_type == "initial time":
            fh.write(f"Time:\n{args[0]}\n")
        elif log_type == "time":
            fh.write(f"{args[0]}\n")
        else:
            fh.write(mesh_info(args[0]) + "\n")
