// UNBUILT SOURCE: this image has no Node toolchain (node, npm, node_api.h are absent).  It shows the N-API
// addon a maintainer of the reference would add to bind include/gsplat_b200.h; see INTEGRATION.md.
// bindings/node/gsplat_napi.cc  —  node-gyp: link -lgsplat_b200, include ../../include
#include <napi.h>
#include "gsplat_b200.h"

class Splats : public Napi::ObjectWrap<Splats> {
 public:
  static Napi::Object Init(Napi::Env env, Napi::Object exports) {
    exports.Set("Splats", DefineClass(env, "Splats", {
      InstanceMethod("clear", &Splats::Clear), InstanceMethod("push", &Splats::Push),
      InstanceMethod("reserve", &Splats::Reserve),
      InstanceMethod("sort", &Splats::Sort),   InstanceMethod("render", &Splats::Render)}));
    return exports;
  }
  explicit Splats(const Napi::CallbackInfo& i) : Napi::ObjectWrap<Splats>(i) {
    int dev = i.Length() ? i[0].As<Napi::Number>().Int32Value() : 0;
    if (gs_create(dev, &ctx_) != GS_OK) Napi::Error::New(i.Env(), gs_last_error(nullptr)).ThrowAsJavaScriptException();
  }
  ~Splats() { gs_destroy(ctx_); }
 private:
  void Check(Napi::Env e, int rc) { if (rc != GS_OK) Napi::Error::New(e, gs_last_error(ctx_)).ThrowAsJavaScriptException(); }
  Napi::Value Clear(const Napi::CallbackInfo& i) { Check(i.Env(), gs_clear(ctx_)); return i.Env().Undefined(); }
  // reserve(numVertexes)                           <- initGL(numVertexes), index.js:248-251
  Napi::Value Reserve(const Napi::CallbackInfo& i) {
    Check(i.Env(), gs_reserve(ctx_, i[0].As<Napi::Number>().Uint32Value()));
    return i.Env().Undefined();
  }
  // push(ArrayBuffer rows, vertexCount)            <- pushDataBuffer(buffer, vertexCount), index.js:328
  Napi::Value Push(const Napi::CallbackInfo& i) {
    auto buf = i[0].As<Napi::ArrayBuffer>();
    Check(i.Env(), gs_push_splats(ctx_, buf.Data(), i[1].As<Napi::Number>().Uint32Value()));
    return i.Env().Undefined();
  }
  // sort(Float32Array view, Float32Array|undefined cutout) -> Uint32Array   <- worker "sort", index.js:587-596
  Napi::Value Sort(const Napi::CallbackInfo& i) {
    auto view = i[0].As<Napi::Float32Array>();
    const float* cut = i[1].IsUndefined() ? nullptr : i[1].As<Napi::Float32Array>().Data();
    uint32_t n = 0, cnt = 0; gs_num_splats(ctx_, &n);
    auto out = Napi::Uint32Array::New(i.Env(), n);
    Check(i.Env(), gs_sort(ctx_, view.Data(), cut, out.Data(), &cnt));
    return Napi::Uint32Array::New(i.Env(), cnt, out.ArrayBuffer(), 0);
  }
  // render({proj, modelview, width, height, focal, cutout?, bg?, depth?: Float32Array}, Uint8Array out)  <- onBeforeRender + draw
  Napi::Value Render(const Napi::CallbackInfo& i) {
    auto o = i[0].As<Napi::Object>();
    gs_render_params p{};
    memcpy(p.proj, o.Get("proj").As<Napi::Float32Array>().Data(), 64);
    memcpy(p.modelview, o.Get("modelview").As<Napi::Float32Array>().Data(), 64);
    p.width = o.Get("width").As<Napi::Number>().Uint32Value();
    p.height = o.Get("height").As<Napi::Number>().Uint32Value();
    p.focal = o.Get("focal").As<Napi::Number>().FloatValue();
    if (o.Has("cutout")) { p.has_cutout = 1; memcpy(p.cutout16, o.Get("cutout").As<Napi::Float32Array>().Data(), 64); }
    if (o.Has("depth")) p.depth_in = o.Get("depth").As<Napi::Float32Array>().Data();  // gl.readPixels(DEPTH) of the scene so far
    p.out_format = GS_FORMAT_RGBA8;
    Check(i.Env(), gs_render(ctx_, &p, i[1].As<Napi::Uint8Array>().Data(), nullptr));
    return i.Env().Undefined();
  }
  gs_context* ctx_ = nullptr;
};
Napi::Object InitAll(Napi::Env env, Napi::Object exports) { return Splats::Init(env, exports); }
NODE_API_MODULE(gsplat_b200, InitAll)
