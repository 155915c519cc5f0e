"""CLI mirrors of steps 12 / 13 / 14: argument parsing and the .npz cube container on CPU; the
full 12 -> 13 -> 14 chain against the oracle chain on the GPU."""
import os

import numpy as np
import pytest
import yaml

from pseudo_3d_interpolation_b200 import cube_io, synth
from pseudo_3d_interpolation_b200 import cube_apply_FFT, cube_apply_IFFT, cube_POCS_interpolation_3D as step13


def _time_cube(nt=64, n1=24, n2=20, seed=3):
    rng = np.random.default_rng(seed)
    ev = synth.draw_events(rng, nt, n1, n2, n_events=4)
    d, twt = synth.time_cube(ev, nt, n1, n2)
    fold = synth.make_fold(rng, n1, n2, 0.5)
    fold[0, 0] = 3                                   # exercises mask = min(fold, 1)
    sparse = (d * (fold > 0)[None]).astype(np.float32)
    cube = cube_io.Cube(attrs=dict(long_name="test cube", description="synthetic", history="", text=""))
    cube.variables["env"] = (("twt", "iline", "xline"), sparse)
    cube.variables["fold"] = (("iline", "xline"), fold)
    cube.coords = dict(twt=twt, iline=np.arange(n1), xline=np.arange(n2))
    cube.coord_attrs["twt"] = dict(units="ms")
    return cube, d


def test_argparse_matches_reference_flags():
    a = cube_apply_FFT.define_input_args().parse_args(["c.nc", "--params_netcdf", "p.yml", "--compute_real", "--filter", "lowpass",
                                                       "--filter_freqs", "1000", "2000", "--drop-filtered-freq", "-V"])
    assert a.compute_real and a.filter == "lowpass" and a.filter_freqs == [1000, 2000] and a.drop_filtered_freq and a.verbose == 1
    assert a.prefix == "freq" and a.upsampling_factor == 1
    b = step13.define_input_args().parse_args(["c.nc", "--path_pocs_parameter", "cfg.yml", "--verbose", "2"])
    assert b.path_output_dir is None and b.verbose == 2
    c = cube_apply_IFFT.define_input_args().parse_args(["c.nc", "--params_netcdf", "p.yml", "--rescale-envelope"])
    assert c.rescale_envelope and not c.compute_real
    with pytest.raises(SystemExit):
        step13.define_input_args().parse_args(["c.nc"])          # --path_pocs_parameter is required


def test_npz_cube_roundtrip(tmp_path):
    cube, _ = _time_cube()
    cube.variables["z"] = (("twt", "iline", "xline"), (cube.data("env") * (1 + 2j)).astype(np.complex64))
    p = str(tmp_path / "cube_twt.npz")
    cube_io.write_cube(p, cube, split_complex=True)
    back = cube_io.open_cube(p)
    assert set(back.data_vars) == {"env", "fold", "z.real", "z.imag"}
    assert back.dims_of("env") == ("twt", "iline", "xline") and back.other_dim() == "twt"
    np.testing.assert_array_equal(back.data("fold"), cube.data("fold"))
    np.testing.assert_array_equal(back.data("z.imag"), cube.data("z").imag)
    assert back.coord_attrs["twt"]["units"] == "ms" and back.attrs["long_name"] == "test cube"


def test_netcdf_needs_xarray(tmp_path):
    if cube_io._have_xarray():
        pytest.skip("xarray present")
    with pytest.raises(ImportError):
        cube_io.open_cube(str(tmp_path / "x.nc"))


@pytest.mark.gpu
@pytest.mark.parametrize("compute_real", [True, False])
def test_steps_12_13_14_chain_matches_oracle(tmp_path, compute_real):
    from oracle import pocs_oracle as orc, time_axis_oracle as tor
    cube, dense = _time_cube()
    p_time = str(tmp_path / "cube_twt.npz")
    cube_io.write_cube(p_time, cube)
    p_nc = str(tmp_path / "attrs.yml")
    with open(p_nc, "w") as f:
        yaml.safe_dump(dict(attrs_time=dict(env=dict(units="amp"), twt=dict(units="ms")), attrs_freq=dict(data=dict(units="-"), new_dim=dict(units="kHz"))), f)
    meta = dict(transform_kind="fft", niter=15, eps=0.0, thresh_op="soft", thresh_model="exponential", alpha=1.0, p_max=0.99,
                p_min=1e-3, sqrt_decay=False, decay_kind="values", version="regular")
    p_cfg = str(tmp_path / "pocs.yml")
    with open(p_cfg, "w") as f:
        yaml.safe_dump(dict(dim="freq_twt", var="freq_env", batch_chunk=20, n_workers=4, processes=True, threads_per_worker=1,
                            memory_limit="2GB", output_runtime_results=True, metadata=meta), f)
    flag = ["--compute_real"] if compute_real else []
    # step 12
    fcube = cube_apply_FFT.main(["12", p_time, "--params_netcdf", p_nc] + flag, return_dataset=True)
    p_freq = str(tmp_path / "cube_freq.npz")
    assert os.path.exists(p_freq)
    F_ref, f_ref = tor.time_fft(cube.data("env"), cube.coords["twt"], compute_real=compute_real)
    assert np.linalg.norm(fcube.data("freq_env") - F_ref) / np.linalg.norm(F_ref) < 1e-5
    assert fcube.var_attrs["freq_env"]["original_var"] == "env"
    # step 13
    res = step13.main(["13", p_freq, "--path_pocs_parameter", p_cfg], return_dataset=True)
    out_dir = str(tmp_path / "cube_freq_FFT_soft_niter-15")
    assert os.path.isdir(out_dir) and os.path.exists(os.path.join(out_dir, "parameter_cube_freq_FFT_soft_niter-15.yml"))
    assert os.path.exists(os.path.join(out_dir, "runtimes_cube_freq_FFT_soft_niter-15.txt"))
    assert os.path.exists(out_dir + ".npz")
    stored = cube_io.open_cube(out_dir + ".npz")
    assert {"freq_env_interp.real", "freq_env_interp.imag", "fold"} <= set(stored.data_vars)
    assert np.all(np.diff(res.coords["freq_twt"]) > 0)                     # merged axis ascending
    params = {k: v for k, v in meta.items() if k != "transform_kind"}
    y_ref = orc.pocs_cube(F_ref, cube.data("fold"), **params)
    order = np.argsort(f_ref, kind="stable")
    got = res.data("freq_env_interp")
    assert np.linalg.norm(got - y_ref[order]) / np.linalg.norm(y_ref) < 1e-4
    # step 14
    tcube = cube_apply_IFFT.main(["14", out_dir + ".npz", "--params_netcdf", p_nc] + flag, return_dataset=True)
    assert os.path.exists(str(tmp_path / "cube_twt_FFT_soft_niter-15_interp-freq.npz"))
    x_ref = tor.time_ifft(y_ref[order], synth.DT_MS, synth.T0_MS, compute_real=compute_real, ascending=True)
    x = tcube.data("env")
    assert x.dtype == np.float32 and x.shape == x_ref.shape
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 2e-4
    np.testing.assert_allclose(tcube.coords["twt"], cube.coords["twt"].astype(np.float32), rtol=1e-6)
    # the interpolation actually helps: closer to the dense cube than the sparse input is
    sparse = cube.data("env")
    assert np.linalg.norm(x - dense) < 0.8 * np.linalg.norm(sparse - dense)


@pytest.mark.gpu
@pytest.mark.parametrize("nt,compute_real", [(64, True), (64, False), (512, True), (2048, True)])
def test_on_device_chain_matches_oracle_chain(nt, compute_real):
    """12 -> 13 -> 14 chained on the device (pipeline.interpolate_time_cube) against the oracle
    chain time_fft -> pocs_cube -> time_ifft; nt = 512 / 2048 exercise the register-resident
    transposing time-axis kernels, nt = 64 the generic ones."""
    from oracle import pocs_oracle as orc, time_axis_oracle as tor
    from pseudo_3d_interpolation_b200 import pipeline
    n1, n2 = (20, 18) if nt >= 512 else (24, 20)
    cube, dense = _time_cube(nt=nt, n1=n1, n2=n2, seed=8)
    x, twt, fold = cube.data("env"), cube.coords["twt"], cube.data("fold")
    params = dict(niter=12, thresh_op="soft", thresh_model="linear", eps=0.0, alpha=1.0, p_max=0.99, p_min=1e-2)
    res = {}
    y, spec = pipeline.interpolate_time_cube(x, twt, fold, compute_real=compute_real, return_spectrum=True, results=res, **params)
    F_ref, f = tor.time_fft(x, twt, compute_real=compute_real)
    S_ref = orc.pocs_cube(F_ref, fold, **params)
    x_ref = tor.time_ifft(S_ref, synth.DT_MS, synth.T0_MS, compute_real=compute_real, ascending=False)
    assert y.shape == x_ref.shape and y.dtype == np.float32
    assert np.linalg.norm(spec - S_ref) / np.linalg.norm(S_ref) < 1e-4
    assert np.linalg.norm(y - x_ref) / np.linalg.norm(x_ref) < 1e-4
    assert np.all(res["niterations"][np.abs(F_ref).reshape(F_ref.shape[0], -1).max(axis=1) > 0] == 12)


@pytest.mark.gpu
def test_fused_cli_equals_three_step_chain(tmp_path):
    """`13_cube_interpolate_POCS cube_twt --fused`: steps 12 -> 13 -> 14 chained on the device give the same time cube as the
    three scripts with their intermediate files, and the YAML `precision` key is honoured."""
    cube, _ = _time_cube(nt=128, n1=30, n2=26, seed=11)
    p_time = str(tmp_path / "cube_twt.npz")
    cube_io.write_cube(p_time, cube)
    p_nc = str(tmp_path / "attrs.yml")
    with open(p_nc, "w") as f:
        yaml.safe_dump(dict(attrs_time=dict(env=dict(units="amp"), twt=dict(units="ms")), attrs_freq=dict(data=dict(units="-"), new_dim=dict(units="kHz"))), f)
    meta = dict(transform_kind="fft", niter=14, eps=0.0, thresh_op="hard", thresh_model="exponential", alpha=1.0, p_max=0.99, p_min=1e-4)
    for precision in ("auto", 64):
        p_cfg = str(tmp_path / f"pocs_{precision}.yml")
        with open(p_cfg, "w") as f:
            yaml.safe_dump(dict(dim="freq_twt", var="freq_env", precision=precision, metadata=meta), f)
        cube_apply_FFT.main(["12", p_time, "--params_netcdf", p_nc, "--compute_real"], return_dataset=True)
        step13.main(["13", str(tmp_path / "cube_freq.npz"), "--path_pocs_parameter", p_cfg], return_dataset=True)
        three = cube_apply_IFFT.main(["14", str(tmp_path / "cube_freq_FFT_hard_niter-14.npz"), "--params_netcdf", p_nc, "--compute_real"],
                                     return_dataset=True).data("env")
        with open(p_cfg, "w") as f:
            yaml.safe_dump(dict(var="env", precision=precision, metadata=meta), f)
        fused = step13.main(["13", p_time, "--path_pocs_parameter", p_cfg, "--fused", "--compute_real"], return_dataset=True)
        assert os.path.exists(str(tmp_path / "cube_twt_interp.npz"))
        y = fused.data("env_interp")
        assert y.dtype == np.float32 and y.shape == three.shape
        assert np.linalg.norm(y - three) / np.linalg.norm(three) < 1e-5
