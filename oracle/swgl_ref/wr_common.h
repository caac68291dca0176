// TEST INFRASTRUCTURE — not product code.
//
// Shared scaffolding for the hand-instantiated SWGL shader programs that let
// the UNMODIFIED reference rasteriser (/root/reference/swgl/src/gl.cc and its
// headers) be built here without the Rust `glsl-to-cxx` translator.  Upstream
// generates one `<name>_program` C++ class per GLSL program at build time
// (swgl/build.rs:60-112, glsl-to-cxx/src/lib.rs:195-245); this file provides
// the pieces those generated classes share, written by hand in the `glsl.h`
// vocabulary, following the `#ifdef SWGL` branches of the GLSL in
// webrender/res/.  Nothing here contains raster, blend or sampling
// arithmetic: that all comes from the reference headers at compile time.
//
// This header is included from load_shader.h, i.e. in the middle of gl.cc
// (gl.cc:2663), so `using namespace glsl` and every SWGL global is in scope.

#pragma once

// ---------------------------------------------------------------------------
// Uniform locations.  Fixed for every program so the driver can be simple.
// (Upstream numbers them per program; the host always looks them up by name
// through GetUniformLocation, device/gl.rs:3085, so any numbering works.)
enum WrUniform {
  WR_U_sColor0 = 1,
  WR_U_sColor1,
  WR_U_sColor2,
  WR_U_sGpuCache,
  WR_U_sTransformPalette,
  WR_U_sRenderTasks,
  WR_U_sDither,
  WR_U_sPrimitiveHeadersF,
  WR_U_sPrimitiveHeadersI,
  WR_U_sClipMask,
  WR_U_sGpuBufferF,
  WR_U_sGpuBufferI,
  WR_U_uTransform,
  WR_U_uMode,
};

static inline int wr_uniform_index(const char* name) {
  static const struct {
    const char* n;
    int i;
  } table[] = {
      {"sColor0", WR_U_sColor0},
      {"sColor1", WR_U_sColor1},
      {"sColor2", WR_U_sColor2},
      {"sGpuCache", WR_U_sGpuCache},
      {"sTransformPalette", WR_U_sTransformPalette},
      {"sRenderTasks", WR_U_sRenderTasks},
      {"sDither", WR_U_sDither},
      {"sPrimitiveHeadersF", WR_U_sPrimitiveHeadersF},
      {"sPrimitiveHeadersI", WR_U_sPrimitiveHeadersI},
      {"sClipMask", WR_U_sClipMask},
      {"sGpuBufferF", WR_U_sGpuBufferF},
      {"sGpuBufferI", WR_U_sGpuBufferI},
      {"uTransform", WR_U_uTransform},
      {"uMode", WR_U_uMode},
  };
  for (auto& e : table) {
    if (strcmp(e.n, name) == 0) return e.i;
  }
  return -1;
}

// Which samplers a program declares (so init_batch only prepares those, as the
// generated bind_textures() would; glsl-to-cxx/src/lib.rs:332-352).
enum WrSamplerBit {
  WR_S_Color0 = 1 << 0,
  WR_S_Color1 = 1 << 1,
  WR_S_Color2 = 1 << 2,
  WR_S_GpuCache = 1 << 3,
  WR_S_TransformPalette = 1 << 4,
  WR_S_RenderTasks = 1 << 5,
  WR_S_Dither = 1 << 6,
  WR_S_PrimitiveHeadersF = 1 << 7,
  WR_S_PrimitiveHeadersI = 1 << 8,
  WR_S_ClipMask = 1 << 9,
  WR_S_GpuBufferF = 1 << 10,
  WR_S_GpuBufferI = 1 << 11,
};

// Named attribute table (glsl-to-cxx/src/lib.rs:436-465 emits one struct per
// program; a small dynamic table is equivalent).
struct WrAttribs {
  static const int MAX = 16;
  const char* names[MAX] = {};
  int locs[MAX];
  int count = 0;
  int add(const char* name) {
    names[count] = name;
    locs[count] = NULL_ATTRIB;
    return count++;
  }
  void bind_loc(const char* name, int index) {
    for (int i = 0; i < count; i++) {
      if (strcmp(names[i], name) == 0) {
        locs[i] = index;
        return;
      }
    }
  }
  int get_loc(const char* name) const {
    for (int i = 0; i < count; i++) {
      if (strcmp(names[i], name) == 0) {
        return locs[i] != NULL_ATTRIB ? locs[i] : -1;
      }
    }
    return -1;
  }
};

struct WrCommon {
  struct Samplers {
    sampler2D_impl sColor0_impl, sColor1_impl, sColor2_impl, sGpuCache_impl,
        sTransformPalette_impl, sRenderTasks_impl, sDither_impl,
        sPrimitiveHeadersF_impl, sClipMask_impl, sGpuBufferF_impl;
    isampler2D_impl sPrimitiveHeadersI_impl, sGpuBufferI_impl;
    int slot[16] = {};
    bool set_slot(int index, int value) {
      if (index >= WR_U_sColor0 && index <= WR_U_sGpuBufferI) {
        slot[index] = value;
        return true;
      }
      return false;
    }
  } samplers;
  WrAttribs attrib_locations;
  unsigned sampler_mask = 0;

  sampler2D sColor0 = nullptr, sColor1 = nullptr, sColor2 = nullptr,
            sGpuCache = nullptr, sTransformPalette = nullptr,
            sRenderTasks = nullptr, sDither = nullptr,
            sPrimitiveHeadersF = nullptr, sClipMask = nullptr,
            sGpuBufferF = nullptr;
  isampler2D sPrimitiveHeadersI = nullptr, sGpuBufferI = nullptr;
  mat4_scalar uTransform;
  int uMode = 0;

  void bind_textures() {
#define WR_BIND(bit, name, fn)                                        \
  if (sampler_mask & bit)                                             \
    name = fn(&samplers.name##_impl, samplers.slot[WR_U_##name]);
    WR_BIND(WR_S_Color0, sColor0, lookup_sampler)
    WR_BIND(WR_S_Color1, sColor1, lookup_sampler)
    WR_BIND(WR_S_Color2, sColor2, lookup_sampler)
    WR_BIND(WR_S_GpuCache, sGpuCache, lookup_sampler)
    WR_BIND(WR_S_TransformPalette, sTransformPalette, lookup_sampler)
    WR_BIND(WR_S_RenderTasks, sRenderTasks, lookup_sampler)
    WR_BIND(WR_S_Dither, sDither, lookup_sampler)
    WR_BIND(WR_S_PrimitiveHeadersF, sPrimitiveHeadersF, lookup_sampler)
    WR_BIND(WR_S_PrimitiveHeadersI, sPrimitiveHeadersI, lookup_isampler)
    WR_BIND(WR_S_ClipMask, sClipMask, lookup_sampler)
    WR_BIND(WR_S_GpuBufferF, sGpuBufferF, lookup_sampler)
    WR_BIND(WR_S_GpuBufferI, sGpuBufferI, lookup_isampler)
#undef WR_BIND
  }

  // ---- shared.glsl:77  get_fetch_uv(i, vpi) -------------------------------
  static ivec2_scalar get_fetch_uv(int i, uint32_t vpi) {
    return ivec2_scalar(int(vpi * (uint32_t(i) % (1024U / vpi))),
                        int(uint32_t(i) / (1024U / vpi)));
  }
  // ---- gpu_cache.glsl:16-19 / gpu_buffer.glsl:13-16 -----------------------
  static ivec2_scalar get_gpu_cache_uv(int address) {
    return ivec2_scalar(int(uint32_t(address) % 1024U),
                        int(uint32_t(address) / 1024U));
  }
  vec4_scalar fetch_from_gpu_cache_1(int address) {
    return texelFetch(sGpuCache, get_gpu_cache_uv(address), 0);
  }
  vec4_scalar fetch_gpu_cache(int address, int offset) {
    ivec2_scalar uv = get_gpu_cache_uv(address);
    return texelFetch(sGpuCache, uv + ivec2_scalar(offset, 0), 0);
  }
  vec4_scalar fetch_gpu_buffer_f(int address, int offset) {
    ivec2_scalar uv = get_gpu_cache_uv(address);
    return texelFetch(sGpuBufferF, uv + ivec2_scalar(offset, 0), 0);
  }
  ivec4_scalar fetch_from_gpu_buffer_1i(int address) {
    return texelFetch(sGpuBufferI, get_gpu_cache_uv(address), 0);
  }

  // ---- rect.glsl ----------------------------------------------------------
  struct RectWithEndpoint {
    vec2_scalar p0;
    vec2_scalar p1;
  };

  // ---- transform.glsl:22-46 ----------------------------------------------
  struct Transform {
    mat4_scalar m;
    mat4_scalar inv_m;
    bool is_axis_aligned;
  };
  Transform fetch_transform(int id) {
    Transform transform;
    transform.is_axis_aligned = (id >> 23) == 0;
    int index = id & 0x007fffff;
    ivec2_scalar uv0 = get_fetch_uv(index, 8U);
    for (int i = 0; i < 4; i++) {
      transform.m[i] =
          texelFetch(sTransformPalette, uv0 + ivec2_scalar(i, 0), 0);
      transform.inv_m[i] =
          texelFetch(sTransformPalette, uv0 + ivec2_scalar(4 + i, 0), 0);
    }
    return transform;
  }

  // ---- render_task.glsl ---------------------------------------------------
  struct RenderTaskData {
    RectWithEndpoint task_rect;
    vec4_scalar user_data;
  };
  RenderTaskData fetch_render_task_data(int index) {
    ivec2_scalar uv = get_fetch_uv(index, 2U);
    vec4_scalar texel0 = texelFetch(sRenderTasks, uv + ivec2_scalar(0, 0), 0);
    vec4_scalar texel1 = texelFetch(sRenderTasks, uv + ivec2_scalar(1, 0), 0);
    RenderTaskData data;
    data.task_rect = RectWithEndpoint{texel0.sel(X, Y), texel0.sel(Z, W)};
    data.user_data = texel1;
    return data;
  }
  struct PictureTask {
    RectWithEndpoint task_rect;
    float device_pixel_scale;
    vec2_scalar content_origin;
  };
  PictureTask fetch_picture_task(int address) {
    RenderTaskData task_data = fetch_render_task_data(address);
    return PictureTask{task_data.task_rect, task_data.user_data.x,
                       task_data.user_data.sel(Y, Z)};
  }
  struct ClipArea {
    RectWithEndpoint task_rect;
    float device_pixel_scale;
    vec2_scalar screen_origin;
  };
  ClipArea fetch_clip_area(int index) {
    RenderTaskData task_data;
    if (index >= 0x7FFFFFFF) {
      task_data.task_rect =
          RectWithEndpoint{vec2_scalar(0.0f), vec2_scalar(0.0f)};
      task_data.user_data = vec4_scalar(0.0f);
    } else {
      task_data = fetch_render_task_data(index);
    }
    return ClipArea{task_data.task_rect, task_data.user_data.x,
                    task_data.user_data.sel(Y, Z)};
  }
};

// Boilerplate the translator emits for every vertex shader
// (glsl-to-cxx/src/lib.rs:354-434, 3601-3648).
#define WR_VERTEX_ABI(NAME)                                                    \
  static void set_uniform_1i(VertexShaderImpl* impl, int index, int value) {   \
    Self* self = (Self*)impl;                                                  \
    if (self->samplers.set_slot(index, value)) return;                         \
    if (index == WR_U_uMode) self->uMode = value;                              \
  }                                                                            \
  static void set_uniform_4fv(VertexShaderImpl*, int, const float*) {}         \
  static void set_uniform_matrix4fv(VertexShaderImpl* impl, int index,         \
                                    const float* value) {                      \
    Self* self = (Self*)impl;                                                  \
    if (index == WR_U_uTransform)                                              \
      self->uTransform = mat4_scalar::load_from_ptr(value);                    \
  }                                                                            \
  static void run(VertexShaderImpl* impl, char* interps,                       \
                  size_t interp_stride) {                                      \
    Self* self = (Self*)impl;                                                  \
    self->main();                                                              \
    self->store_interp_outputs(interps, interp_stride);                        \
  }                                                                            \
  static void init_batch(VertexShaderImpl* impl) {                             \
    Self* self = (Self*)impl;                                                  \
    self->bind_textures();                                                     \
  }                                                                            \
  void init_vertex_abi() {                                                     \
    this->set_uniform_1i_func = &set_uniform_1i;                                     \
    this->set_uniform_4fv_func = &set_uniform_4fv;                                   \
    this->set_uniform_matrix4fv_func = &set_uniform_matrix4fv;                       \
    this->init_batch_func = &init_batch;                                             \
    this->load_attribs_func = &Self::load_attribs;                                         \
    this->run_primitive_func = &run;                                                 \
  }

// Boilerplate for the fragment side (lib.rs:3563-3636).
#define WR_FRAGMENT_ABI()                                                      \
  static void run(FragmentShaderImpl* impl) {                                  \
    Self* self = (Self*)impl;                                                  \
    self->main();                                                              \
    self->step_interp_inputs();                                                \
  }                                                                            \
  static void skip(FragmentShaderImpl* impl, int steps) {                      \
    Self* self = (Self*)impl;                                                  \
    self->step_interp_inputs(steps);                                           \
  }                                                                            \
  void init_fragment_abi() {                                                   \
    this->init_span_func = &read_interp_inputs;                                      \
    this->run_func = &run;                                                           \
    this->skip_func = &skip;                                                         \
    this->init_span_w_func = &read_interp_inputs;                                    \
    this->run_w_func = &run;                                                         \
    this->skip_w_func = &skip;                                                       \
  }

// Fragment shaders with varyings (or reading gl_FragCoord.z/.w) also get the perspective variants the
// translator emits (lib.rs:660-690, 716-745, 3576-3590, 3627-3631): the shader defines
// read_perspective_inputs / step_perspective_inputs next to the plain pair and uses this macro instead.
#define WR_FRAGMENT_ABI_W()                                                    \
  static void run(FragmentShaderImpl* impl) {                                  \
    Self* self = (Self*)impl;                                                  \
    self->main();                                                              \
    self->step_interp_inputs();                                                \
  }                                                                            \
  static void skip(FragmentShaderImpl* impl, int steps) {                      \
    Self* self = (Self*)impl;                                                  \
    self->step_interp_inputs(steps);                                           \
  }                                                                            \
  static void run_perspective(FragmentShaderImpl* impl) {                      \
    Self* self = (Self*)impl;                                                  \
    self->main();                                                              \
    self->step_perspective_inputs();                                           \
  }                                                                            \
  static void skip_perspective(FragmentShaderImpl* impl, int steps) {          \
    Self* self = (Self*)impl;                                                  \
    self->step_perspective_inputs(steps);                                      \
  }                                                                            \
  void init_fragment_abi() {                                                   \
    this->init_span_func = &read_interp_inputs;                                \
    this->run_func = &run;                                                     \
    this->skip_func = &skip;                                                   \
    this->enable_perspective();                                                \
    this->init_span_w_func = &read_perspective_inputs;                         \
    this->run_w_func = &run_perspective;                                       \
    this->skip_w_func = &skip_perspective;                                     \
  }

// Boilerplate for the program class (lib.rs:224-241).
#define WR_PROGRAM(NAME, KEY)                                                  \
  struct NAME##_program : ProgramImpl, NAME##_frag {                           \
    int get_uniform(const char* name) const override {                         \
      return wr_uniform_index(name);                                           \
    }                                                                          \
    void bind_attrib(const char* name, int index) override {                   \
      attrib_locations.bind_loc(name, index);                                  \
    }                                                                          \
    int get_attrib(const char* name) const override {                          \
      return attrib_locations.get_loc(name);                                   \
    }                                                                          \
    size_t interpolants_size() const override {                                \
      return sizeof(InterpOutputs);                                            \
    }                                                                          \
    VertexShaderImpl* get_vertex_shader() override { return this; }            \
    FragmentShaderImpl* get_fragment_shader() override { return this; }        \
    const char* get_name() const override { return KEY; }                      \
    static ProgramImpl* loader() { return new NAME##_program; }                \
  };
