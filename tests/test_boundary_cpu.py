"""Drop-in boundary tests (CPU): the argv GPUStack composes for this engine parses with our CLI,
the backend definitions satisfy GPUStack's registration rules, and the placeholder / parameter
flattening semantics match the reference's own table tests
(tests/worker/backends/test_backend.py:78-152 in the reference)."""
from pathlib import Path

import pytest
import yaml

import gpustack_compat as gc   # tests/gpustack_compat.py (test infrastructure)
from gpustack_b200.cli import build_parser

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("params,expected", [
    (["--ctx-size 1024"], ["--ctx-size", "1024"]),
    ([" --ctx-size=1024"], ["--ctx-size=1024"]),
    (["--ctx-size =1024"], ["--ctx-size=1024"]),
    (["--ctx-size = 1024"], ["--ctx-size=1024"]),
    (["--max-model-len 8192", "--enforce-eager"], ["--max-model-len", "8192", "--enforce-eager"]),
    (['--served-model-name "my model"'], ["--served-model-name", "my model"]),
    (["--foo=bar=baz"], ["--foo=bar=baz"]),
    (None, []),
])
def test_flatten_backend_param_table(params, expected):
    assert gc.flatten_backend_param(params) == expected


def test_replace_command_param_placeholders_and_env():
    cmd = "serve {{model_path}} --host {{worker_ip}} --port {{port}} --served-model-name {{model_name}} --x {{FOO}} --y {{MISSING}}"
    out = gc.replace_command_param(cmd, "/models/llama", 40001, "10.0.0.7", "llama3", env={"FOO": "1"})
    assert out == "serve /models/llama --host 10.0.0.7 --port 40001 --served-model-name llama3 --x 1 --y {{MISSING}}"
    assert gc.replace_command_param("", "/m", 1) == ""


def test_custom_backend_yaml_is_valid_and_cli_accepts_its_argv():
    backend = yaml.safe_load((ROOT / "deploy" / "inference-backend.yaml").read_text())
    assert gc.validate_backend(backend) == []
    argv = gc.build_command_args(backend, None, "/models/llama3-8b", 40003, "192.168.1.5", "llama3-8b",
                                 backend_parameters=["--max-model-len 4096", "--max-num-seqs=128"])
    assert argv[0] == "b200-engine"
    args, unknown = build_parser().parse_known_args(argv[1:])
    assert unknown == []
    assert (args.cmd, args.model, args.host, args.port) == ("serve", "/models/llama3-8b", "192.168.1.5", 40003)
    assert args.served_model_name == ["llama3-8b"] and args.max_model_len == 4096 and args.max_num_seqs == 128
    assert args.tensor_parallel_size == "auto"      # nothing injects TP for custom backends


def test_vllm_custom_version_argv_is_accepted_with_injected_flags():
    spec = yaml.safe_load((ROOT / "deploy" / "vllm-custom-version.yaml").read_text())
    vcfg = spec["version_configs"]["b200-custom"]
    argv = gc.build_vllm_command_args(vcfg, "/models/l3", 40010, "10.1.1.1", "l3", gpu_count=8,
                                      derived_max_model_len=131072,
                                      backend_parameters=["--gpu-memory-utilization 0.85", "--no-enable-prefix-caching",
                                                          "--enable-expert-parallel"])
    assert argv[:3] == ["b200-engine", "serve", "/models/l3"]
    assert argv[3:7] == ["--max-model-len", "8192", "--tensor-parallel-size", "8"]
    args, unknown = build_parser().parse_known_args(argv[1:])
    assert args.tensor_parallel_size == "8" and args.max_model_len == 8192 and args.port == 40010
    assert args.gpu_memory_utilization == 0.85
    assert args.enable_prefix_caching is False
    assert unknown == ["--enable-expert-parallel"]    # tolerated (warned about), not fatal
    # user-set parallelism / host are respected, not duplicated
    argv = gc.build_vllm_command_args(vcfg, "/m", 1, "h", "n", gpu_count=4, backend_parameters=["-tp 2", "--host 0.0.0.0"])
    assert "--tensor-parallel-size" not in argv and argv.count("--host") == 1


def test_registration_rules():
    bad = {"backend_name": "b200-native", "version_configs": {"v0": {"image_name": "x"}}}
    errs = gc.validate_backend(bad)
    assert any("-custom" in e for e in errs) and any("custom_framework" in e for e in errs)
    assert any("health_check_path" in e for e in errs)


def test_first_class_backend_binding_compiles():
    """deploy/first_class_backend/b200_native.py is the reference-side binding a GPUStack maintainer
    would add (BackendEnum / _SERVER_CLASS_MAPPING / selector switch).  gpustack.* cannot be imported
    offline, so it is compiled, and the engine-side names it calls must exist."""
    import ast

    src = (ROOT / "deploy" / "first_class_backend" / "b200_native.py").read_text()
    tree = ast.parse(src)
    classes = {n.name for n in tree.body if isinstance(n, ast.ClassDef)}
    assert {"B200NativeServer", "B200NativeResourceFitSelector"} <= classes
    from gpustack_b200.config import tiny_llama, EngineConfig
    from gpustack_b200.resource_fit import estimate

    est = estimate(tiny_llama(), EngineConfig(tensor_parallel_size=2))
    assert est.total_min == est.per_gpu_min * 2 and est.total_min > 0
