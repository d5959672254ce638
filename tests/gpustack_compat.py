"""Host-side mirror of how GPUStack's worker composes this engine's command line.

GPUStack cannot be imported offline (gpustack-runtime / sqlmodel are not installed), so the
semantics of the launcher functions on the path are restated here for the boundary tests and as
executable documentation of INTEGRATION.md.  Same names and argument meaning as the reference:

  replace_command_param   gpustack/schemas/inference_backend.py:176-222
  flatten_backend_param   gpustack/worker/backends/base.py:867-891
  build_command_args      gpustack/worker/backends/custom.py:141-159            (custom backend)
  build_vllm_command_args gpustack/worker/backends/vllm.py:395-507,602-639      (custom vLLM version)
  validate_backend        gpustack/routes/inference_backend.py:1226-1255        (-custom suffix rules)
"""
from __future__ import annotations

import re
import shlex
from typing import Dict, List, Optional, Sequence

_VAR = re.compile(r"\{\{([A-Za-z_][A-Za-z0-9_]*)\}\}")


def replace_command_param(command: str, model_path: Optional[str], port: Optional[int],
                          worker_ip: Optional[str] = None, model_name: Optional[str] = None,
                          env: Optional[Dict[str, str]] = None) -> str:
    """{{model_path}}, {{port}}, {{worker_ip}}, {{model_name}} then {{ENV_VAR}} from `env`;
    unknown variables are left in place."""
    if not command:
        return ""
    command = command.replace("{{model_path}}", model_path or "")
    command = command.replace("{{port}}", str(port))
    command = command.replace("{{worker_ip}}", worker_ip or "")
    command = command.replace("{{model_name}}", model_name or "")
    if env:
        command = _VAR.sub(lambda m: env.get(m.group(1), m.group(0)), command)
    return command


def flatten_backend_param(backend_parameters: Optional[Sequence[str]]) -> List[str]:
    """["--ctx-size 1024"] -> ["--ctx-size", "1024"]; ["--k = v"] / ["--k =v"] -> ["--k=v"]."""
    out: List[str] = []
    for param in backend_parameters or []:
        p = param.strip()
        if "=" in p:
            k, v = (x.strip() for x in p.split("=", 1))
            out.append(f"{k}={v}")
        else:
            out.extend(shlex.split(p))
    return out


def build_command_args(backend: dict, version: Optional[str], model_path: str, port: int, worker_ip: str,
                       model_name: str, backend_parameters: Optional[Sequence[str]] = None,
                       env: Optional[Dict[str, str]] = None, run_command: Optional[str] = None) -> List[str]:
    """argv (entrypoint + args) GPUStack's CustomServer would start for this backend definition."""
    version = version or backend.get("default_version")
    vcfg = (backend.get("version_configs") or {}).get(version) or {}
    command = run_command or vcfg.get("run_command") or backend.get("default_run_command") or ""
    inline = replace_command_param(command, model_path, port, worker_ip, model_name, env)
    args = shlex.split(inline) if inline else []
    args += flatten_backend_param(list(backend.get("default_backend_param") or []) + list(backend_parameters or []))
    entry = vcfg.get("entrypoint") or backend.get("default_entrypoint")
    return (shlex.split(entry) if entry else []) + args


def find_parameter(params: Optional[Sequence[str]], names: Sequence[str]) -> Optional[str]:
    """Value of the first parameter whose name (leading dashes stripped) is in `names`; accepts
    "--k=v", "--k v" inside one string, and "--k", "v" as two tokens
    (gpustack/utils/command.py:22-50)."""
    if params is None:
        return None
    for i, p in enumerate(params):
        p = p.strip()
        if "=" in p:
            k, v = p.split("=", 1)
            if k.strip().lstrip("-") in names:
                return v
        elif " " in p:
            k, v = p.split(" ", 1)
            if k.strip().lstrip("-") in names:
                parts = shlex.split(v)
                return parts[0] if len(parts) == 1 else v
        elif p.lstrip("-") in names:
            if i + 1 < len(params):
                return params[i + 1]
    return None


def build_vllm_command_args(version_cfg: dict, model_path: str, port: int, worker_ip: str, model_name: str,
                            gpu_count: int, derived_max_model_len: Optional[int] = None,
                            backend_parameters: Optional[Sequence[str]] = None) -> List[str]:
    """argv GPUStack's VLLMServer composes for a custom version of the built-in vLLM backend."""
    user = flatten_backend_param(backend_parameters)
    args = shlex.split(replace_command_param(version_cfg.get("run_command") or "{{model_path}}", model_path, port,
                                             worker_ip, model_name))
    if derived_max_model_len and derived_max_model_len > 8192 and find_parameter(user, ["max-model-len"]) is None:
        args += ["--max-model-len", "8192"]
    parallel = find_parameter(user, ["tensor-parallel-size", "tp", "pipeline-parallel-size", "pp",
                                     "data-parallel-size", "dp"])
    if gpu_count > 1 and parallel is None:
        args += ["--tensor-parallel-size", str(gpu_count)]
    args += user
    for flag, val in (("host", worker_ip), ("port", str(port)), ("served-model-name", model_name)):
        if find_parameter(user, [flag]) is None:
            args += [f"--{flag}", val]
    entry = version_cfg.get("entrypoint")
    return (shlex.split(entry) if entry else ["vllm", "serve"]) + args


def validate_backend(backend: dict) -> List[str]:
    """Problems GPUStack would reject (or that would make the engine misbehave in GPUStack)."""
    errs = []
    name = backend.get("backend_name", "")
    builtin = name in ("vLLM", "SGLang", "MindIE", "VoxBox")
    if not builtin and not name.endswith("-custom"):
        errs.append("backend_name must end with '-custom'")
    for ver, cfg in (backend.get("version_configs") or {}).items():
        if not ver.endswith("-custom"):
            errs.append(f"version '{ver}' must end with '-custom'")
        if not cfg.get("image_name"):
            errs.append(f"version '{ver}' needs image_name")
        if not cfg.get("custom_framework"):
            errs.append(f"version '{ver}' needs custom_framework (e.g. cuda)")
    if not builtin and not backend.get("health_check_path"):
        errs.append("health_check_path should be set: a custom backend without one is always 'ready'")
    return errs
