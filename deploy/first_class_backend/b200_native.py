"""Reference-side binding for registering this engine as a FIRST-CLASS GPUStack backend (SURVEY.md
section 8f-3) instead of a `-custom` backend: the file a GPUStack maintainer would add as
`gpustack/worker/backends/b200_native.py`, plus the three one-line hooks listed at the bottom.

It is written against the reference's own interfaces and imports them (gpustack.* and
gpustack_runtime are not installable offline, so this file is syntax-checked here, not imported:
tests/test_boundary_cpu.py::test_first_class_backend_binding_compiles).  Nothing in
gpustack_b200/ depends on it -- the engine stays a process + CLI + HTTP drop-in.

What the first-class path adds over the custom-backend YAML (deploy/inference-backend.yaml):
  * `--tensor-parallel-size N` injected from the scheduled GPU count, exactly like the built-in vLLM
    backend (gpustack/worker/backends/vllm.py:602-639 `get_auto_parallelism_arguments`);
  * `--max-model-len 8192` capping when the checkpoint derives more and the user set nothing
    (vllm.py:427-436);
  * the engine's EXACT VRAM requirement as the resource claim (gpustack_b200.resource_fit: weights
    per dtype + workspace + CUDA-graph pool + the KV capacity asked for) instead of the
    `weights x 1.2 + 2 GiB` heuristic of CustomBackendResourceFitSelector
    (gpustack/policies/utils.py:283-360);
  * the model's `speculative_config` (n-gram method) and `extended_kv_cache` turned into the engine's
    `--speculative-config` / `--kv-host-cache-gib` flags, which the built-in vLLM backend does for its own engine
    (vllm.py:256-275,333-367,468-483) and a custom backend cannot (it only sees backend_parameters).
"""
import json
from typing import List

from gpustack.policies.candidate_selectors.custom_backend_resource_fit_selector import (
    CustomBackendResourceFitSelector,
)
from gpustack.utils.command import find_parameter
from gpustack.worker.backends.custom import CustomServer
from gpustack.worker.backends.vllm import get_auto_parallelism_arguments


class B200NativeServer(CustomServer):
    """Container lifecycle is CustomServer's (create_workload with image / GPUs / mounts / ports,
    gpustack/worker/backends/custom.py:47-135); only the argv differs."""

    def _build_command_args(self) -> List[str]:
        port = self._get_serving_port()
        args = ["serve", self._model_path]
        args = self.build_versioned_command_args(args)          # base.py:606-662 (custom versions / entrypoints)
        if find_parameter(self._model.backend_parameters, ["max-model-len"]) is None:
            derived = self._derive_max_model_len()
            if derived and derived > 8192:
                args += ["--max-model-len", "8192"]
        # single worker, several GPUs -> --tensor-parallel-size len(gpu_indexes); multi-worker is not
        # supported by this engine (one NVSwitch node), so is_distributed is always False
        args += get_auto_parallelism_arguments(self._model.backend_parameters, self._model_instance, False)
        args += self._speculative_arguments() + self._extended_kv_arguments()
        args += self._flatten_backend_param()                    # base.py:867-891
        for flag, value in (("--host", self._worker.ip), ("--port", str(port)),
                            ("--served-model-name", self._model.name)):
            if find_parameter(args, [flag.lstrip("-")]) is None:
                args += [flag, value]
        return args


    def _speculative_arguments(self) -> List[str]:
        """Model.speculative_config -> --speculative-config JSON (the field names the vLLM launcher uses, so the
        engine's parser is the same for both registration routes).  Only the n-gram method has an implementation
        in the engine; eagle3 / mtp models are launched without speculation (the engine would warn and ignore)."""
        sc = getattr(self._model, "speculative_config", None)
        if not sc or not sc.enabled or str(getattr(sc.algorithm, "value", sc.algorithm)).lower() != "ngram":
            return []
        cfg = {"method": "ngram"}
        if sc.num_draft_tokens:
            cfg["num_speculative_tokens"] = sc.num_draft_tokens
        if sc.ngram_max_match_length:
            cfg["prompt_lookup_max"] = sc.ngram_max_match_length
        if sc.ngram_min_match_length:
            cfg["prompt_lookup_min"] = sc.ngram_min_match_length
        return ["--speculative-config", json.dumps(cfg)]

    def _extended_kv_arguments(self) -> List[str]:
        """Model.extended_kv_cache -> the host-memory tier of the prefix cache, sized like the reference sizes
        LMCache: explicit ram_size (GiB), else ram_ratio x the VRAM claim."""
        kv = getattr(self._model, "extended_kv_cache", None)
        if not kv or not kv.enabled:
            return []
        if kv.ram_size and kv.ram_size > 0:
            gib = float(kv.ram_size)
        else:
            claim = sum((self._model_instance.computed_resource_claim.vram or {}).values()) if \
                self._model_instance.computed_resource_claim else 0
            gib = claim * float(kv.ram_ratio or 1.2) / (1 << 30)
        return ["--kv-host-cache-gib", f"{gib:.2f}"] if gib > 0 else []


class B200NativeResourceFitSelector(CustomBackendResourceFitSelector):
    """Same placement logic as the custom selector (manual gpu_ids / gpus_per_replica, or all
    allocatable GPUs of a worker), with the VRAM claim computed by the engine's own model
    (`b200-engine estimate <model_dir> -tp N` prints the same numbers)."""

    def _set_vram_claim(self):
        from gpustack_b200.config import EngineConfig, ModelConfig
        from gpustack_b200.resource_fit import estimate

        tp = max(1, int(self._gpu_count or 1))
        mcfg = ModelConfig.from_pretrained(self._model.local_path or self._model_path)
        ecfg = EngineConfig(model=str(self._model.local_path), tensor_parallel_size=tp,
                            max_model_len=int(find_parameter(self._model.backend_parameters, ["max-model-len"]) or 0) or None)
        self._vram_claim = estimate(mcfg, ecfg).total_min      # = per-GPU minimum x tp; the selector splits it evenly (:289-292)


# ---------------------------------------------------------------------------------------------------
# The three hooks (one line each) that complete the registration:
#
#   gpustack/schemas/models.py:57-62            class BackendEnum:   B200_NATIVE = "B200Native"
#   gpustack/worker/serve_manager.py:58-63      _SERVER_CLASS_MAPPING[BackendEnum.B200_NATIVE] = B200NativeServer
#   gpustack/scheduler/scheduler.py:416-446     elif model.backend == BackendEnum.B200_NATIVE:
#                                                   candidates_selector = B200NativeResourceFitSelector(config, model, model_instances)
#
# plus a built-in InferenceBackend row (gpustack/schemas/inference_backend.py:54-95) with
# backend_name "B200Native", the image of deploy/Dockerfile, entrypoint `b200-engine`,
# health_check_path `/health`, and the `vLLM` runtime metrics mapping (the engine exports vllm:* names,
# gpustack/assets/metrics_config/metrics_config.yaml:74-98).
