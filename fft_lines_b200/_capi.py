"""ctypes view of libfftlines_cuda.so (the C-ABI in include/*.h).

Loading never falls back to anything: a missing library raises, and on a
machine without a GPU every compute entry point returns FFTL_ERR_CUDA.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFTL_LIB") or os.path.join(HERE, "libfftlines_cuda.so")  # FFTL_LIB: developer A/B builds

FFTL_OK, FFTL_ERR_ARG, FFTL_ERR_CUDA, FFTL_ERR_NOMEM, FFTL_ERR_STATE = 0, 1, 2, 3, 4
WINDOW_RECT, WINDOW_HANN = 0, 1
BINNING_LINEAR, BINNING_LOG = 0, 1
BEAT_SAMPLES, TEMPO_RING = 160, 256
MOVE_EXACT, MOVE_REFERENCE = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FAST_REGISTERS = 0, 1, 2
TICK_AUTO, TICK_THREE_LAUNCHES = 0, 1


class FftlConfig(C.Structure):
    """fftl_config (include/fftlines.h)."""
    _fields_ = [
        ("n", C.c_int32), ("hop", C.c_int32), ("bars", C.c_int32), ("window", C.c_int32),
        ("binning", C.c_int32), ("sg_half", C.c_int32), ("sg_degree", C.c_int32),
        ("circle", C.c_int32), ("zoom", C.c_float), ("sample_rate", C.c_float),
        ("fmin", C.c_float), ("fmax", C.c_float), ("beat_bars", C.c_int32 * 4),
        ("bass_bin", C.c_int32), ("strict_int", C.c_int32), ("reserved", C.c_int32 * 3),
    ]

    def copy(self):
        c = FftlConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(FftlConfig))
        return c


class FftlSettings(C.Structure):
    """fftl_settings: the 11 numbers of the reference's Settings.txt (settingsFile.c:376)."""
    _fields_ = [("bar_count", C.c_int32), ("zoom", C.c_float), ("color_sel", C.c_int32), ("dofft", C.c_int32),
                ("background", C.c_int32), ("gradient", C.c_int32), ("ignore_serial", C.c_int32), ("circle", C.c_int32),
                ("waveform", C.c_int32), ("stereo", C.c_int32), ("beat_detection", C.c_int32)]


class FftlBeat(C.Structure):
    """fftl_beat (include/fftlines.h)."""
    _fields_ = [
        ("w", C.c_int32), ("thr", C.c_int32), ("beat_value", C.c_int32), ("flag", C.c_int32),
        ("bass", C.c_int32), ("peak_index", C.c_int32), ("movement_per_sec", C.c_float),
        ("reserved", C.c_int32),
    ]


# Every symbol include/*.h declares, by header (used by the CPU-side export test).
EXPORTED_FUNCTIONS = {
    "fftlines.h": [
        "fftl_config_reference", "fftl_config_extended", "fftl_config_validate", "fftl_build_bins",
        "fftl_build_sg", "fftl_stft_create", "fftl_stft_destroy", "fftl_stft_num_frames", "fftl_stft_first_frame_needed",
        "fftl_stft_sample_span", "fftl_stft_run_device", "fftl_stft_run_device_span", "fftl_stft_run", "fftl_stft_launch_count",
        "fftl_stft_fast_launch_count", "fftl_stft_tmem_launch_count", "fftl_stft_set_kernel", "fftl_stft_timing",
        "fftl_stream_create", "fftl_stream_destroy", "fftl_stream_reset", "fftl_stream_push_device",
        "fftl_stream_push", "fftl_stream_push_device_f32", "fftl_stream_launch_count", "fftl_stream_set_tick_kernel", "fftl_stream_fused_tick_count", "fftl_stream_set_history", "fftl_stream_history_device",
        "fftl_stream_waveform_device", "fftl_settings_default", "fftl_settings_parse", "fftl_settings_format",
        "fftl_settings_to_config", "fftl_color_index", "fftl_led_bytes_device", "fftl_marker_sweep_device",
        "fftl_host_alloc", "fftl_host_free", "fftl_synth_pcm_device", "fftl_measure_fp32_peak", "fftl_last_error",
        "fftl_last_error_string", "fftl_clear_error", "fftl_version", "fftl_device_count",
        "fftl_set_n", "fftl_get_n", "fftl_compat_reset_beat", "fftl_compat_launch_count",
    ],
    "fft_calculate.h": ["initializefft", "executefft", "cleanupfft", "initializefft256", "executefft256",
                        "cleanupfft256"],
    "beatDetector.h": ["AddBeatSample", "GetBeatValue", "BassBeatDetector"],
}
EXPORTED_DATA = {"beatDetector.h": ["beatMovementPerSec", "beatposition", "timeDelta", "direction"]}

_lib = None


class FftLinesError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libfftlines_cuda error %d: %s" % (code, message))
        self.code = code


def lib():
    """Load (once) and type the library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            "%s is missing: build it with `python -m fft_lines_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    cfgp = C.POINTER(FftlConfig)
    i32p, f32p, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int16)
    vp = C.c_void_p
    L.fftl_config_reference.argtypes = [cfgp, C.c_int32, C.c_int32, C.c_int32]
    L.fftl_config_reference.restype = None
    L.fftl_config_extended.argtypes = [cfgp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float]
    L.fftl_config_extended.restype = None
    L.fftl_config_validate.argtypes = [cfgp]
    L.fftl_build_bins.argtypes = [cfgp, i32p, i32p]
    L.fftl_build_sg.argtypes = [C.c_int32, C.c_int32, f32p]
    L.fftl_stft_create.argtypes = [cfgp, C.c_int, C.POINTER(vp)]
    L.fftl_stft_destroy.argtypes = [vp]
    L.fftl_stft_destroy.restype = None
    L.fftl_stft_num_frames.argtypes = [vp, C.c_int64]
    L.fftl_stft_num_frames.restype = C.c_int64
    L.fftl_stft_first_frame_needed.argtypes = [C.c_int64, C.c_int32]
    L.fftl_stft_first_frame_needed.restype = C.c_int64
    L.fftl_stft_run_device.argtypes = [vp, vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                       vp, vp, vp, vp]
    L.fftl_stft_sample_span.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.fftl_stft_run_device_span.argtypes = [vp, vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                            C.c_int64, vp, vp, vp, vp]
    L.fftl_stft_fast_launch_count.argtypes = [vp]
    L.fftl_stft_fast_launch_count.restype = C.c_int64
    L.fftl_stft_tmem_launch_count.argtypes = [vp]
    L.fftl_stft_tmem_launch_count.restype = C.c_int64
    L.fftl_stft_set_kernel.argtypes = [vp, C.c_int32]
    L.fftl_stft_run.argtypes = [vp, vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp]
    L.fftl_stft_launch_count.argtypes = [vp]
    L.fftl_stft_launch_count.restype = C.c_int64
    L.fftl_stft_timing.argtypes = [vp, i32p, f32p, f32p]
    L.fftl_stream_create.argtypes = [cfgp, C.c_int32, C.c_int32, C.c_int, C.POINTER(vp)]
    L.fftl_stream_destroy.argtypes = [vp]
    L.fftl_stream_destroy.restype = None
    L.fftl_stream_reset.argtypes = [vp]
    L.fftl_stream_push_device.argtypes = [vp, vp, C.c_int32, C.c_int64, vp, vp, vp, vp]
    L.fftl_stream_push.argtypes = [vp, vp, C.c_int32, C.c_int64, vp, vp, vp]
    L.fftl_stream_push_device_f32.argtypes = [vp, vp, C.c_int32, C.c_int64, C.c_int64, vp, vp, vp, vp]
    L.fftl_stream_set_history.argtypes = [vp, C.c_int32]
    L.fftl_stream_history_device.argtypes = [vp, vp, vp]
    L.fftl_stream_waveform_device.argtypes = [vp, C.c_int32, C.c_float, C.c_float, vp, vp]
    L.fftl_stream_launch_count.argtypes = [vp]
    L.fftl_stream_launch_count.restype = C.c_int64
    L.fftl_stream_fused_tick_count.argtypes = [vp]
    L.fftl_stream_fused_tick_count.restype = C.c_int64
    L.fftl_stream_set_tick_kernel.argtypes = [vp, C.c_int32]
    L.fftl_stream_set_tick_kernel.restype = C.c_int
    L.fftl_settings_default.argtypes = [C.POINTER(FftlSettings)]
    L.fftl_settings_default.restype = None
    L.fftl_settings_parse.argtypes = [C.c_char_p, C.c_int64, C.POINTER(FftlSettings)]
    L.fftl_settings_format.argtypes = [C.POINTER(FftlSettings), C.c_char_p, C.c_int64]
    L.fftl_settings_to_config.argtypes = [C.POINTER(FftlSettings), C.c_int32, C.c_int32, C.c_float, cfgp]
    L.fftl_settings_to_config.restype = None
    L.fftl_color_index.argtypes = [C.c_int32, C.c_int32]
    L.fftl_color_index.restype = C.c_uint32
    L.fftl_led_bytes_device.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp]
    L.fftl_marker_sweep_device.argtypes = [vp, C.c_int64, C.c_float, C.c_int32, C.c_int32, vp, vp, vp]
    L.fftl_host_alloc.argtypes = [C.c_uint64]
    L.fftl_host_alloc.restype = vp
    L.fftl_host_free.argtypes = [vp]
    L.fftl_host_free.restype = None
    L.fftl_synth_pcm_device.argtypes = [vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_uint32, vp]
    L.fftl_measure_fp32_peak.argtypes = [C.POINTER(C.c_double), C.c_int32]
    L.fftl_last_error_string.restype = C.c_char_p
    L.fftl_clear_error.restype = None
    L.fftl_set_n.argtypes = [C.c_int32]
    L.fftl_compat_reset_beat.restype = None
    L.fftl_compat_launch_count.restype = C.c_int64
    for name in ("initializefft", "executefft", "initializefft256", "executefft256"):
        getattr(L, name).argtypes = [vp, vp]
        getattr(L, name).restype = None
    L.cleanupfft.restype = None
    L.cleanupfft256.restype = None
    L.AddBeatSample.argtypes = [C.c_int]
    L.AddBeatSample.restype = None
    L.GetBeatValue.restype = C.c_int
    L.BassBeatDetector.argtypes = [vp, vp, vp, vp]
    L.BassBeatDetector.restype = None
    _lib = L
    return L


def check(rc):
    if rc != FFTL_OK:
        L = lib()
        raise FftLinesError(rc, (L.fftl_last_error_string() or b"").decode())


def last_error():
    L = lib()
    return L.fftl_last_error(), (L.fftl_last_error_string() or b"").decode()


def data_symbol(name, ctype):
    """The reference's exported globals (beatDetector.h:8-11)."""
    return ctype.in_dll(lib(), name)
