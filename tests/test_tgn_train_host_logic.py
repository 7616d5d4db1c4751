"""Host-side logic of the TGN training path on a CPU-only box: the Python face of
tgm_b200/nn/tgn.py (handle life cycle, in-place parameter refresh, autograd routing, argument order
and buffer shapes of tgm_tgn_forward_saved / tgm_tgn_backward / tgm_gae_backward) driven through a
stand-in library built on the numpy oracle (tests/_fake_tgn_lib.py).  The bodies are the GPU tests
of tests/test_gpu_tgn_train.py with DEV = 'cpu'; the CUDA kernels themselves are NOT exercised
here."""
import glob
import os

import numpy as np
import pytest
import torch

from tests import _fake_tgn_lib
from tests import test_gpu_tgn_train as gpu_tests
from tests._golden import GOLDEN_DIR


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(gpu_tests, 'DEV', 'cpu')
    with _fake_tgn_lib.installed() as lib:
        yield lib


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgngrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_memory_gradients_reach_every_parameter(fake, path):
    gpu_tests.test_tgn_memory_gradients_match_reference_autograd(path)
    assert fake.calls['tgm_tgn_forward_saved'] >= 3 and \
        fake.calls['tgm_tgn_backward'] == fake.calls['tgm_tgn_forward_saved']
    assert fake.calls['tgm_tgn_create'] == 1 and 'tgm_tgn_set_params' not in fake.calls


def test_parameter_refresh_is_in_place(fake):
    gpu_tests.test_tgn_memory_parameter_refresh_keeps_state_and_message_stores()
    assert fake.calls['tgm_tgn_create'] == 1 and fake.calls['tgm_tgn_set_params'] == 3


def test_nodes_without_messages(fake):
    gpu_tests.test_tgn_memory_gradients_vs_oracle_with_nodes_without_messages()


@pytest.mark.parametrize('dims', [(40, 300, 5, 100, 7, 2), (33, 65, 8, 6, 0, 3), (50, 0, 16, 8, 4, 4)],
                         ids=['reference_test_dims', 'no_msg_feats', 'no_edges'])
def test_embedding_gradients_reach_every_parameter_and_x(fake, dims):
    gpu_tests.test_graph_attention_embedding_gradients_vs_oracle(dims)
    assert fake.calls['tgm_gae_backward'] == 1


def test_training_step_chains_embedding_into_memory(fake):
    gpu_tests.test_tgn_training_step_memory_into_embedding()
    assert fake.calls['tgm_gae_set_params'] == 1 and fake.calls['tgm_tgn_set_params'] == 1
    assert fake.calls['tgm_gae_create'] == 1 and fake.calls['tgm_tgn_create'] == 1


def test_inference_paths_do_not_record_autograd(fake):
    from tgm_b200.nn import GraphAttentionEmbedding, TGNMemory
    mem = TGNMemory(20, 3, 4, 5).train()
    mem.reset_state()
    n_id = torch.arange(6)
    with torch.no_grad():
        z, _ = mem(n_id)
    assert not z.requires_grad and fake.calls.get('tgm_tgn_forward') == 1
    z, _ = mem.eval()(n_id)          # eval mode: stored rows, no graph (tgn.py:160-161)
    assert not z.requires_grad and 'tgm_tgn_forward_saved' not in fake.calls
    z, _ = mem.train()(torch.zeros(0, dtype=torch.int64))  # empty n_id
    assert z.shape == (0, 4)
    enc = GraphAttentionEmbedding(4, 6, 3, mem.time_enc).eval()
    out = enc(torch.randn(6, 4), torch.zeros(6, dtype=torch.int64), torch.zeros(2, 0, dtype=torch.int64),
              torch.zeros(0, dtype=torch.int64), torch.zeros(0, 3))
    assert not out.requires_grad and 'tgm_gae_backward' not in fake.calls
    # dropout 0.1 (the constructor's default) in training mode: runs with dropout disabled and
    # says so once (the reference's example trains the default-constructed module)
    from tgm_b200.nn import attention as _att
    _att._DROPOUT_WARNED.discard('GraphAttentionEmbedding')
    with pytest.warns(UserWarning, match='dropout p=0.1 is not applied'):
        out = enc.train()(torch.randn(6, 4), torch.zeros(6, dtype=torch.int64),
                          torch.zeros(2, 0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64),
                          torch.zeros(0, 3))
    assert out.shape == (6, 6)


def test_state_dict_interchanges_with_the_reference_checkpoint_layout(fake):
    """The reference registers memory / last_update / _assoc as buffers (tgn.py:128-133), so its
    checkpoints carry them; here they live behind the native handle and are mapped in and out."""
    from tgm_b200.nn import TGNMemory
    N, D, M, TD = 10, 3, 4, 5
    m = TGNMemory(N, D, M, TD)
    sd = m.state_dict()
    assert {'memory', 'last_update', '_assoc', 'memory_updater.weight_ih', 'time_enc.w.weight'} <= set(sd)
    assert sd['memory'].shape == (N, M) and not sd['memory'].any() and sd['last_update'].dtype == torch.int64
    ckpt = {k: v.clone() for k, v in sd.items()}
    ckpt['memory'], ckpt['last_update'] = torch.arange(40.).view(N, M), torch.arange(N)
    m2 = TGNMemory(N, D, M, TD)
    m2.load_state_dict(ckpt)                               # before any handle exists: kept pending
    assert torch.equal(m2.state_dict()['memory'], ckpt['memory']) and 'tgm_tgn_create' not in fake.calls
    assert torch.equal(m2.memory, ckpt['memory']) and torch.equal(m2.last_update, ckpt['last_update'])
    assert '_pending_state' not in m2.__dict__ and fake.calls['tgm_tgn_create'] == 1
    live = m2.state_dict()
    live['memory'] = live['memory'] * 2
    m2.load_state_dict(live)                               # live handle: written through
    assert torch.equal(m2.memory, ckpt['memory'] * 2) and fake.calls['tgm_tgn_create'] == 1
    params_only = {k: v for k, v in sd.items() if k not in ('memory', 'last_update', '_assoc')}
    TGNMemory(N, D, M, TD).load_state_dict(params_only)    # parameter-only checkpoints still load
    bad = dict(sd, memory=torch.zeros(3, M))
    with pytest.raises(RuntimeError, match='memory'):
        TGNMemory(N, D, M, TD).load_state_dict(bad)


# --- MeanAggregator: the same host-side checks for tests/test_gpu_tgn_mean.py ---------------------
from tests import test_gpu_tgn_mean as gpu_mean_tests  # noqa: E402


@pytest.fixture
def fake_mean(monkeypatch):
    monkeypatch.setattr(gpu_mean_tests, 'DEV', 'cpu')
    with _fake_tgn_lib.installed() as lib:
        yield lib


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgnmean_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_mean_aggregator_state_machine_through_the_python_face(fake_mean, path):
    gpu_mean_tests.test_tgn_mean_memory_matches_reference_fixture(path)
    assert fake_mean.calls['tgm_tgn_set_aggregator'] == 1


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgnmeangrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[12:-4])
def test_mean_aggregator_gradients_through_the_python_face(fake_mean, path):
    gpu_mean_tests.test_tgn_mean_memory_gradients_match_reference_autograd(path)
    assert fake_mean.calls['tgm_tgn_backward'] >= 3


def test_mean_aggregator_many_messages_per_node(fake_mean):
    gpu_mean_tests.test_tgn_mean_gradients_vs_oracle_at_c4_dims_with_many_messages_per_node()


def test_unknown_aggregator_is_refused():
    from tgm_b200.nn import TGNMemory
    with pytest.raises(NotImplementedError):
        TGNMemory(10, 3, 4, 5, aggregator_module=torch.nn.Identity())


def test_mean_aggregator_long_run_and_flush(fake_mean):
    gpu_mean_tests.test_tgn_mean_event_log_grows_and_restarts_after_a_flush()
