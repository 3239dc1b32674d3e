// ref_shim/CL/cl2.hpp -- stand-in for the OpenCL header the reference includes
// (BoyreMoore.cpp:4), so that the UNMODIFIED BoyreMoore.cpp builds with g++ on a box without an
// OpenCL SDK.  Buffers are host memory, the "device" is the calling thread, and
// clEnqueueNDRangeKernel runs the reference's own kernel1.cl (compiled as C++) once per global
// id.  Only the calls BoyreMoore.cpp:217-312 makes are provided.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>      // the reference relies on MSVC pulling in clock()/strcpy transitively

typedef int cl_int;
typedef unsigned cl_uint;
typedef unsigned long cl_bitfield;
typedef struct shim_obj { void *host; size_t bytes; } *cl_mem;
typedef void *cl_platform_id, *cl_device_id, *cl_context, *cl_command_queue, *cl_program, *cl_event;
typedef struct shim_kernel { void *arg[8]; int iarg[8]; } *cl_kernel;
#define CL_TRUE 1
#define CL_DEVICE_TYPE_GPU 4
#define CL_MEM_READ_ONLY 4
#define CL_MEM_READ_WRITE 1

static int shim_gid = 0;
static inline int get_global_id(int) { return shim_gid; }
#define __kernel static
#define __global
#include "/root/reference/BoyreMoore/x64/Debug/kernel1.cl"
#undef __kernel
#undef __global

static inline cl_int clGetPlatformIDs(cl_uint, cl_platform_id *p, cl_uint *) { *p = (void *)1; return 0; }
static inline cl_int clGetDeviceIDs(cl_platform_id, cl_bitfield, cl_uint, cl_device_id *d, cl_uint *) { *d = (void *)1; return 0; }
static inline cl_context clCreateContext(void *, cl_uint, cl_device_id *, void *, void *, cl_int *r) { if (r) *r = 0; return (void *)1; }
static inline cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_bitfield, cl_int *r) { if (r) *r = 0; return (void *)1; }
static inline cl_mem clCreateBuffer(cl_context, cl_bitfield, size_t bytes, void *, cl_int *r)
{
	cl_mem m = (cl_mem)malloc(sizeof(shim_obj));
	m->host = calloc(bytes ? bytes : 1, 1); m->bytes = bytes;
	if (r) *r = 0;
	return m;
}
static inline cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem m, cl_uint, size_t off, size_t bytes, const void *src, cl_uint, void *, void *)
{ memcpy((char *)m->host + off, src, bytes); return 0; }
static inline cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem m, cl_uint, size_t off, size_t bytes, void *dst, cl_uint, void *, void *)
{ memcpy(dst, (char *)m->host + off, bytes); return 0; }
static inline cl_program clCreateProgramWithSource(cl_context, cl_uint, const char **, const size_t *, cl_int *r) { if (r) *r = 0; return (void *)1; }
static inline cl_int clBuildProgram(cl_program, cl_uint, cl_device_id *, const char *, void *, void *) { return 0; }
static inline cl_kernel clCreateKernel(cl_program, const char *, cl_int *r) { if (r) *r = 0; return (cl_kernel)calloc(1, sizeof(shim_kernel)); }
static inline cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t bytes, const void *val)
{
	if (bytes == sizeof(cl_mem) && idx < 6) k->arg[idx] = (*(cl_mem *)val)->host;   // args 0-5 are buffers
	else k->iarg[idx] = *(const int *)val;                                         // arg 6 is the int length
	return 0;
}
static inline cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel k, cl_uint, const size_t *, const size_t *global, const size_t *, cl_uint, void *, cl_event *)
{
	for (size_t g = 0; g < global[0]; ++g) {
		shim_gid = (int)g;
		search((char *)k->arg[0], (char *)k->arg[1], (int *)k->arg[2], (int *)k->arg[3], (int *)k->arg[4], (int *)k->arg[5], k->iarg[6]);
	}
	return 0;
}
static inline cl_int clFlush(cl_command_queue) { return 0; }
static inline cl_int clReleaseKernel(cl_kernel k) { free(k); return 0; }
static inline cl_int clReleaseProgram(cl_program) { return 0; }
static inline cl_int clReleaseMemObject(cl_mem m) { free(m->host); free(m); return 0; }
static inline cl_int clReleaseCommandQueue(cl_command_queue) { return 0; }
static inline cl_int clReleaseContext(cl_context) { return 0; }
