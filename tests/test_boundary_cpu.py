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


def test_first_class_backend_binding_builds_the_engine_argv_under_stand_in_gpustack_modules(tmp_path):
    """The binding imports gpustack.* (not installable offline).  With stand-ins for exactly the four names it
    imports -- behaving as the reference documents them -- `_build_command_args` must compose the argv our CLI parses:
    TP from the GPU count, max-model-len capping, speculative / extended-KV model fields as engine flags, user
    parameters, host / port / served name.  The argv is then fed to the real `b200-engine` parser."""
    import importlib.util
    import json
    import sys
    import types

    from tests import gpustack_compat as G

    mods = {}

    def mod(name):
        m = types.ModuleType(name)
        mods[name] = m
        return m

    for pkg in ("gpustack", "gpustack.policies", "gpustack.policies.candidate_selectors", "gpustack.utils",
                "gpustack.worker", "gpustack.worker.backends"):
        mod(pkg).__path__ = []
    mod("gpustack.policies.candidate_selectors.custom_backend_resource_fit_selector").CustomBackendResourceFitSelector = object
    mod("gpustack.utils.command").find_parameter = G.find_parameter

    class CustomServer:                                   # the members the binding relies on (custom.py / base.py)
        def _get_serving_port(self):
            return 40123

        def build_versioned_command_args(self, args):
            return list(args)

        def _derive_max_model_len(self):
            return 131072

        def _flatten_backend_param(self):
            return G.flatten_backend_param(self._model.backend_parameters)

    mod("gpustack.worker.backends.custom").CustomServer = CustomServer
    mod("gpustack.worker.backends.vllm").get_auto_parallelism_arguments = (
        lambda params, inst, dist: [] if G.find_parameter(params, ["tensor-parallel-size", "tp"]) else
        ["--tensor-parallel-size", str(len(inst.gpu_indexes))])
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("b200_native_binding",
                                                      ROOT / "deploy" / "first_class_backend" / "b200_native.py")
        binding = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(binding)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    srv = binding.B200NativeServer.__new__(binding.B200NativeServer)
    ns = types.SimpleNamespace
    srv._model_path = "/models/llama3-8b"
    srv._worker = ns(ip="10.0.0.5")
    srv._model = ns(name="llama3-8b", backend_parameters=["--max-num-seqs 128", "--enable-prefix-caching"],
                    speculative_config=ns(enabled=True, algorithm=ns(value="ngram"), num_draft_tokens=3,
                                          ngram_max_match_length=4, ngram_min_match_length=2, draft_model=None),
                    extended_kv_cache=ns(enabled=True, ram_size=None, ram_ratio=0.5, chunk_size=None))
    srv._model_instance = ns(gpu_indexes=[0, 1, 2, 3], computed_resource_claim=ns(vram={i: 40 << 30 for i in range(4)}))
    argv = srv._build_command_args()
    assert argv[:2] == ["serve", "/models/llama3-8b"]
    from gpustack_b200.cli import _engine_config, build_parser

    a = build_parser().parse_args(argv)
    assert (a.tensor_parallel_size, a.max_model_len, a.max_num_seqs, a.host, a.port) == ("4", 8192, 128, "10.0.0.5", 40123)
    assert a.served_model_name == ["llama3-8b"]
    e = _engine_config(a, 4)
    assert e.speculative().num_speculative_tokens == 3 and e.speculative().prompt_lookup_max == 4
    assert abs(e.kv_host_cache_gib - 80.0) < 1e-6                  # 0.5 x (4 x 40 GiB)
    assert json.loads(a.speculative_config)["method"] == "ngram"
    # an eagle3 model: launched without speculation instead of failing the deployment
    srv._model.speculative_config = ns(enabled=True, algorithm=ns(value="eagle3"), num_draft_tokens=2,
                                       ngram_max_match_length=None, ngram_min_match_length=None, draft_model="d")
    srv._model.extended_kv_cache = ns(enabled=False, ram_size=None, ram_ratio=1.2, chunk_size=None)
    srv._model.backend_parameters = ["--max-model-len=4096", "--tensor-parallel-size 2"]
    argv = srv._build_command_args()
    a = build_parser().parse_args(argv)
    assert a.speculative_config is None and a.kv_host_cache_gib is None and a.max_model_len == 4096
    assert a.tensor_parallel_size == "2" and argv.count("--tensor-parallel-size") == 1
    # the selector's claim = the engine's own VRAM model for the scheduled GPU count
    from gpustack_b200.config import EngineConfig, llama3_8b
    from gpustack_b200.resource_fit import estimate

    d = tmp_path / "llama3-8b"
    d.mkdir()
    (d / "config.json").write_text(json.dumps(llama3_8b().to_hf_dict()))
    sel = binding.B200NativeResourceFitSelector.__new__(binding.B200NativeResourceFitSelector)
    sel._gpu_count = 2
    sel._model = ns(local_path=str(d), backend_parameters=["--max-model-len 8192"])
    sel._model_path = str(d)
    sel._set_vram_claim()
    want = estimate(llama3_8b(), EngineConfig(tensor_parallel_size=2, max_model_len=8192)).total_min
    assert sel._vram_claim == want and 16e9 < want < 40e9          # 16 GB of weights + workspace + one 8k sequence of KV
