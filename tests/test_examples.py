"""The switched-over cosmology driver (examples/cosmo_experiment.py) runs end to end through the public API."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location('cosmo_experiment', os.path.join(ROOT, 'examples', 'cosmo_experiment.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_hyper_parameters_follow_the_reference_driver():
    """experiments/cosmo/experiment_deepsphere.py:24-117: layer lists, patch index lists, batch size 8 * order, evaluation
    frequency, learning-rate end points (demo.ipynb:847 prints 2.0e-04 at step 0)."""
    ex = _load()
    p = ex.get_params(ntrain=320, EXP_NAME='x', order=2, Nside=1024, architecture='FCN', verbose=False)
    assert p['F'] == [16, 32, 64, 64, 64, 2] and p['K'] == [5] * 6 and p['batch_size'] == 16 and p['statistics'] == 'mean'
    assert p['nsides'] == [1024, 512, 256, 128, 64, 32, 32]
    assert [len(i) for i in p['indexes']] == [262144, 65536, 16384, 4096, 1024, 256, 256]
    assert p['eval_frequency'] == int(80 * 320 / 16 / 80)
    assert p['scheduler'](0) == pytest.approx(2e-4) and p['scheduler'](1000) == pytest.approx(2e-4 * 0.999 ** 1000)
    c = ex.get_params(ntrain=320, EXP_NAME='x', order=2, Nside=1024, architecture='CNN', verbose=False)
    assert c['M'] == [2] and c['statistics'] is None and len(c['F']) == 5 and len(c['nsides']) == 6
    with pytest.raises(ValueError):
        ex.get_params(320, 'x', 2, 1024, architecture='CNN-2d', verbose=False)


@pytest.mark.gpu
@pytest.mark.parametrize('arch, device_data', [('FCN', False), ('CNN', True)])
def test_cosmo_experiment_runs(arch, device_data):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    ex = _load()
    err, t_batch = ex.single_experiment(order=2, sigma_noise=0.5, experiment_type=arch, new=False, n_neighbors=8, Nside=128,
                                        ntrain=64, ntest=32, num_epochs=2, device_data=device_data, verbose=False)
    assert 0.0 <= err <= 1.0 and t_batch > 0
