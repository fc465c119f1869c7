// cz_err.h — thread-local error message behind cz_last_error().
#pragma once
#include <stdarg.h>
#include <stdio.h>
const char* cz_err_get();
int cz_fail(int code, const char* fmt, ...)
#if defined(__GNUC__)
    __attribute__((format(printf, 2, 3)))
#endif
    ;
