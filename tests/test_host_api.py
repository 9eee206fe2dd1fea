"""CPU tests of the host-only behaviour of the drop-in modules (no kernel launches): the parts of the reference
API that are plain Python there too — bias_model containers (modeling/bias.py:9-122), the dispersion_model
container, its scalar accessors, pickling and JSON wire format (modeling/dispersion.pyx:26-57, 59-167, 473-549),
nbinom's closed forms (stats/distributions/nbinom.pyx:140-174)."""
import json
import pickle

import numpy as np
import pytest

from footprint_tools import synth
from footprint_tools.modeling import bias, dispersion
from footprint_tools.modeling.predict import reverse_complement
from footprint_tools.stats.distributions import nbinom


def test_bias_model_container(tmp_path):
    table = synth.vierstra_table()
    letters = "ACGT"
    path = tmp_path / "model.txt"
    with open(path, "w") as f:
        for i in range(0, 4096, 7):     # a model that lacks most k-mers; lower-case keys are upper-cased on read
            kmer = "".join(letters[(i >> (2 * (5 - j))) & 3] for j in range(6))
            f.write("%s\t%r\n" % (kmer.lower() if i % 2 else kmer, float(table[i])))
    bm = bias.kmer_model(str(path))
    assert bm.k == 6 and bm.mid == 3 and bm.offset() == 3
    assert bm["AAAAAA"] == table[0] and bm["AAAAAC"] == 1e-6 and bm["NNNNNN"] == 1e-6 and bm["AAAANA"] == 1e-6
    t = bm.table()
    assert t.shape == (4096,) and np.array_equal(t[::7], table[::7])
    mask = np.ones(4096, dtype=bool)
    mask[::7] = False
    assert np.all(t[mask] == 1e-6)
    assert bias.kmer_index("ACGTAC") == int("012301", 4) and bias.kmer_index("ACGTNC") == -1
    sh = bm.shuffle()
    assert sorted(sh.model.values()) == sorted(bm.model.values()) and sorted(sh.model) == sorted(bm.model)
    with pytest.raises(IOError):
        bias.kmer_model(str(tmp_path / "missing.txt"))
    u = bias.uniform_model()
    assert len(u.model) == 4096 and u["ACGTAC"] == 1.0 and u["NNNNNN"] == 1e-6
    assert np.array_equal(u.probs("ACGTNNACGT"), np.ones(10))     # one value per character (bias.py:121-122)
    assert bm.predict(np.array([1.0, 1.0, 2.0]), 8).tolist() == [2.0, 2.0, 4.0]


def test_reverse_complement_host():
    assert reverse_complement("ACGTN") == "NACGT"
    assert reverse_complement("aacg") == "cgtt" or reverse_complement("aacg") == "CGTT"
    assert reverse_complement("") == ""


def test_dispersion_model_container_and_wire_format(tmp_path):
    dm = dispersion.dispersion_model()
    dm.mu_params, dm.r_params = list(synth.MU_PARAMS), list(synth.R_PARAMS)
    assert isinstance(dm.mu_params, np.ndarray) and dm.mu_params.flags.c_contiguous
    # scalar accessors: three / five segments, value = intercept + slope * x, floors 0.1 and 1e-6
    x0, x1, _, y0, y1, y2, k0, k1, k2 = synth.MU_PARAMS
    for x, want in ((3.0, y0 + k0 * 3.0), (x0, y1 + k1 * x0), (45.0, y1 + k1 * 45.0), (x1, y2 + k2 * x1), (500.0, y2 + k2 * 500.0)):
        assert dm.fit_mu(x) == want
    bad = dispersion.dispersion_model()
    bad.mu_params = [1, 2, 3, -5.0, -5.0, -5.0, 0, 0, 0]
    bad.r_params = [1, 2, 3, 4, 5, -1.0, -1.0, -1.0, -1.0, -1.0, 0, 0, 0, 0, 0]
    assert bad.fit_mu(0.5) == 0.1 and bad.fit_r(0.5) == 1e-6
    r = synth.R_PARAMS
    assert dm.fit_r(7.0) == 1.0 / (r[6] + r[11] * 7.0) and dm.fit_r(100.0) == 1.0 / (r[9] + r[14] * 100.0)
    assert dm.fit_r(1000.0) == 1e-6      # 1 / (0.12 - 0.5) < 0: floored (dispersion.pyx:160-162)
    assert dispersion.piecewise_three(45.0, *synth.MU_PARAMS) == y1 + k1 * 45.0
    assert dispersion.piecewise_five(np.array([1.0, 7.0]), *synth.R_PARAMS).tolist() == [r[5] + r[10] * 1.0, r[6] + r[11] * 7.0]
    with pytest.raises(NotImplementedError):
        str(dm)
    # pickling keeps the fit parameters only (dispersion.pyx:76-85)
    dm.h = np.arange(6, dtype=np.int64).reshape(2, 3)
    dm.p, dm.r = np.array([0.5, np.nan]), np.array([2.0, np.nan])
    dm2 = pickle.loads(pickle.dumps(dm))
    assert np.array_equal(dm2.mu_params, dm.mu_params) and np.array_equal(dm2.r_params, dm.r_params) and dm2.h is None
    # JSON wire format: [dtype, base64, shape] triples, readable by the reference's loader and back
    text = dispersion.write_dispersion_model(dm, extra="unit test")
    d = json.loads(text)
    assert d["mu_params"][0] == "float64" and d["mu_params"][2] == [9] and d["h"][0] == "int64" and d["h"][2] == [2, 3]
    assert d["metadata"] == "unit test" and d["version"].startswith("footprint_tools ")
    path = tmp_path / "dm.json"
    path.write_text(text)
    back = dispersion.load_dispersion_model(str(path))
    assert np.array_equal(back.mu_params, dm.mu_params) and np.array_equal(back.r_params, dm.r_params)
    assert np.array_equal(back.h, dm.h) and np.array_equal(back.p, dm.p, equal_nan=True) and back.metadata == "unit test"
    enc = dispersion.base64encode(np.array([[1.5, -2.0]]))
    assert enc[0] == "float64" and tuple(enc[2]) == (1, 2) and np.array_equal(dispersion.base64decode(enc), [[1.5, -2.0]])
    with pytest.raises(ValueError):
        dm.p_values(np.zeros((2, 2)), np.zeros((2, 2)))      # typed double[:] in the reference: 1-D only


def test_nbinom_closed_forms():
    assert nbinom.mean(0.25, 4.0) == 0.25 * 4.0 / 0.75
    assert nbinom.var(0.25, 4.0) == (0.25 * 4.0) / (0.75 * 0.75)
    with pytest.raises(NotImplementedError):
        nbinom.rvs(0.5, 2.0)
    rng = np.random.default_rng(1)
    data = rng.negative_binomial(5.0, 5.0 / (5.0 + 12.0), 20000).astype(float)
    p, r = nbinom.fit(data)
    assert abs(r - 5.0) < 0.5 and abs(p * r / (1 - p) - 12.0) < 0.3
    eq = nbinom.mle([p, r], data, data.mean())
    assert np.all(np.abs(eq) < 1e-5 * len(data))
