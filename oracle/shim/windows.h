/* Minimal stand-in for <windows.h> so the reference's UNMODIFIED
 * fft_calculate.c / beatDetector.c compile on Linux (they include it through
 * global.h:1 and settingsFile.h:2 but use only these names).
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef FFTL_SHIM_WINDOWS_H
#define FFTL_SHIM_WINDOWS_H
#include <stdint.h>
#include <stddef.h>
typedef void* HWND;
typedef long HRESULT;
typedef int16_t INT16;
typedef intptr_t LRESULT;
typedef uintptr_t WPARAM;
typedef intptr_t LPARAM;
typedef unsigned int UINT;
typedef union _LARGE_INTEGER { long long QuadPart; } LARGE_INTEGER;
#define E_FAIL ((HRESULT)0x80004005L)
static inline int PostMessageW(HWND h, UINT m, WPARAM w, LPARAM l) { (void)h; (void)m; (void)w; (void)l; return 1; }
#endif
