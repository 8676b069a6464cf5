/*
 * node_api_min.h — the subset of Node-API (N-API, ABI-stable since Node 8/10) that fluid_napi.c
 * uses, declared locally because this image ships neither Node nor node_api.h.  Names, argument
 * order and enum values follow the published node_api.h / js_native_api.h (NAPI_VERSION 8); when
 * building against a real Node toolchain, compile with -DFLUID_USE_SYSTEM_NODE_API to include the
 * official header instead.  COMPILE-CHECKED ONLY in this repository (no Node to load it).
 */
#ifndef NODE_API_MIN_H
#define NODE_API_MIN_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected,
               napi_name_expected, napi_function_expected, napi_number_expected,
               napi_boolean_expected, napi_array_expected, napi_generic_failure,
               napi_pending_exception } napi_status;
typedef enum { napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1,
               napi_configurable = 1 << 2 } napi_property_attributes;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array,
               napi_uint16_array, napi_int32_array, napi_uint32_array, napi_float32_array,
               napi_float64_array } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef struct {
    const char* utf8name; napi_value name; napi_callback method; napi_callback getter;
    napi_callback setter; napi_value value; napi_property_attributes attributes; void* data;
} napi_property_descriptor;

#ifdef __cplusplus
extern "C" {
#endif
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv,
                             napi_value* this_arg, void** data);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count,
                                   const napi_property_descriptor* properties);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char* utf8name, bool* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint,
                                 napi_value* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length,
                                   napi_value arraybuffer, size_t byte_offset, napi_value* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type,
                                     size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_create_external_arraybuffer(napi_env env, void* external_data, size_t byte_length,
                                             napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_arraybuffer_info(napi_env env, napi_value arraybuffer, void** data, size_t* byte_length);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_bool(napi_env env, napi_value value, bool* result);
napi_status napi_is_exception_pending(napi_env env, bool* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
#ifdef __cplusplus
}
#endif

#define NAPI_AUTO_LENGTH SIZE_MAX
#define NAPI_MODULE_INIT() \
    napi_value napi_register_module_v1(napi_env env, napi_value exports)
#endif
