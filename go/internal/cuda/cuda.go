// Package cuda binds libizpi_cuda.so (include/izpi_cuda.h) for Izpi's render loop.
//
// NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no Go toolchain in the image).  It is the cgo stub a
// maintainer adds to flynn-nrg/izpi as internal/cuda; INTEGRATION.md explains where each call goes.
//
// Threading: one goroutine per GPU, pinned with runtime.LockOSThread (CUDA context affinity); calls on
// one Context are serialised by that goroutine.  C never retains a Go pointer: uploads copy, outputs
// are Go-allocated slices that live for the duration of the call.
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../izpi_b200 -lizpi_cuda -Wl,-rpath,${SRCDIR}/../../../izpi_b200
#include "izpi_cuda.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Context is one GPU, or a group of GPUs of one box driven by one host thread per GPU inside the library (NewGroup).
type Context struct{ h *C.izpi_ctx }

func lastError(op string, rc C.int) error {
	return fmt.Errorf("izpi cuda: %s: %s (code %d)", op, C.GoString(C.izpi_last_error()), int(rc))
}

// NewContext opens device `dev`.  Call from the goroutine that will own it.
func NewContext(dev int) (*Context, error) {
	runtime.LockOSThread()
	id := C.int(dev)
	var h *C.izpi_ctx
	if rc := C.izpi_ctx_create(1, &id, &h); rc != 0 {
		return nil, lastError("izpi_ctx_create", rc)
	}
	return &Context{h: h}, nil
}

// NewGroup opens all of `devs` as ONE context: the scene is flattened once and copied device-to-device, ray batches are
// split, tiles are claimed dynamically from one cursor and RenderFinish merges the members' disjoint tiles.  This is the
// drop-in for RendererImpl.Render's worker pool (renderer.go:126-147) on a multi-GPU box.
func NewGroup(devs []int) (*Context, error) {
	runtime.LockOSThread()
	ids := make([]C.int, len(devs))
	for i, d := range devs {
		ids[i] = C.int(d)
	}
	var h *C.izpi_ctx
	if rc := C.izpi_ctx_create(C.int(len(ids)), ptr(ids), &h); rc != 0 {
		return nil, lastError("izpi_ctx_create", rc)
	}
	return &Context{h: h}, nil
}

func (c *Context) Close() { C.izpi_ctx_destroy(c.h); c.h = nil }

// SceneDesc is built by package hitable (Flatten, see INTEGRATION.md) from scene.Scene.
type SceneDesc struct {
	WorldKind  int32
	Nodes      []C.izpi_bvh4_node // hitable.BVH4.Nodes reinterpreted: identical 128-byte layout
	Prims      []C.izpi_prim_rec
	TriAttrs   []C.izpi_tri_attr
	Xforms     []C.izpi_xform
	Lights     []int32
	Materials  []C.izpi_material_spec
	Textures   []C.izpi_texture_spec
	Spectral   []C.izpi_spectral_texture_spec
	Camera     C.izpi_camera
	HasWorld   bool
	pin        runtime.Pinner // pins what the nested pointers of Textures / Spectral refer to (PinDoubles) until Upload returns
}

// PinDoubles pins a Go slice that a nested pointer of the descriptor (izpi_texture_spec.pixels,
// izpi_spectral_texture_spec.wavelengths / values) refers to, and returns that pointer.  The cgo rules allow Go memory handed
// to C to contain Go pointers only when their targets are pinned; Upload unpins everything when it returns.
func (d *SceneDesc) PinDoubles(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	d.pin.Pin(&s[0])
	return (*C.double)(&s[0])
}

// pinned returns the base pointer of s after pinning it in p (nil for an empty slice).
func pinned[T any](p *runtime.Pinner, s []T) *T {
	if len(s) == 0 {
		return nil
	}
	p.Pin(&s[0])
	return &s[0]
}

func ptr[T any](s []T) *T {
	if len(s) == 0 {
		return nil
	}
	return &s[0]
}

// Upload copies the flattened scene to the GPU (replaces nothing in the reference: the CPU path keeps
// its object graph; this is the additional one-off step after transport.ToScene()).
func (c *Context) Upload(d *SceneDesc) error {
	// cd lives in Go memory and holds Go pointers: every one of them is pinned for the duration of the call
	// (runtime.Pinner, Go >= 1.21), which is what cgocheck requires of Go memory passed to C.
	var p runtime.Pinner
	defer p.Unpin()
	defer d.pin.Unpin()
	var cd C.izpi_scene_desc
	cd.world_kind = C.int32_t(d.WorldKind)
	cd.n_nodes, cd.nodes = C.int32_t(len(d.Nodes)), pinned(&p, d.Nodes)
	cd.n_prims, cd.prims, cd.tri_attrs = C.int32_t(len(d.Prims)), pinned(&p, d.Prims), pinned(&p, d.TriAttrs)
	cd.n_xforms, cd.xforms = C.int32_t(len(d.Xforms)), pinned(&p, d.Xforms)
	cd.n_lights, cd.lights = C.int32_t(len(d.Lights)), (*C.int32_t)(unsafe.Pointer(pinned(&p, d.Lights)))
	cd.n_materials, cd.materials = C.int32_t(len(d.Materials)), pinned(&p, d.Materials)
	cd.n_textures, cd.textures = C.int32_t(len(d.Textures)), pinned(&p, d.Textures)
	cd.n_spectral_textures, cd.spectral_textures = C.int32_t(len(d.Spectral)), pinned(&p, d.Spectral)
	cd.camera = d.Camera
	if d.HasWorld {
		cd.dielectric_has_world = 1
	}
	if rc := C.izpi_scene_upload(c.h, &cd); rc != 0 {
		return lastError("izpi_scene_upload", rc)
	}
	return nil
}

// TraceClosest is hitable.Hitable.Hit for a batch: origins/directions are xyz-interleaved float64.
// ids[i] is the index of the hit primitive in the ORIGINAL hitables list (-1 = miss).
func (c *Context) TraceClosest(org, dir []float64, tMin, tMax float64, ids []int32, t []float64) error {
	n := len(ids)
	if len(org) != 3*n || len(dir) != 3*n || len(t) != n {
		return fmt.Errorf("izpi cuda: TraceClosest: slice lengths disagree")
	}
	if n == 0 {
		return nil
	}
	rc := C.izpi_trace_closest(c.h, C.int64_t(n), (*C.double)(&org[0]), (*C.double)(&dir[0]), C.double(tMin), C.double(tMax),
		C.IZPI_TRACE_EXACT, (*C.int32_t)(&ids[0]), (*C.double)(&t[0]), nil)
	if rc != 0 {
		return lastError("izpi_trace_closest", rc)
	}
	return nil
}

// RenderConfig mirrors the arguments of render.New (renderer.go:73-88).
type RenderConfig struct {
	Width, Height, Samples, MaxDepth int
	Spectral                         bool
	Background                       [3]float64
	BgWavelengths, BgValues          []float64
	Seed                             uint64
}

func (c *Context) RenderSetup(rc RenderConfig) error {
	var cc C.izpi_render_config
	cc.width, cc.height, cc.spp, cc.max_depth = C.int32_t(rc.Width), C.int32_t(rc.Height), C.int32_t(rc.Samples), C.int32_t(rc.MaxDepth)
	if rc.Spectral {
		cc.sampler = C.IZPI_SAMPLER_SPECTRAL
	}
	cc.sample_count = cc.spp
	for i := 0; i < 3; i++ {
		cc.background[i] = C.double(rc.Background[i])
	}
	var p runtime.Pinner // cc holds Go pointers: pin their targets for the call
	defer p.Unpin()
	if len(rc.BgWavelengths) > 0 && len(rc.BgValues) == len(rc.BgWavelengths) {
		cc.bg_wavelengths = (*C.double)(pinned(&p, rc.BgWavelengths))
		cc.bg_values = (*C.double)(pinned(&p, rc.BgValues))
		cc.n_bg = C.int32_t(len(rc.BgWavelengths))
	}
	cc.seed = C.uint64_t(rc.Seed)
	if r := C.izpi_render_setup(c.h, &cc); r != 0 {
		return lastError("izpi_render_setup", r)
	}
	return nil
}

// RenderTiles renders a batch of workUnits {x0,y0,x1,y1} (renderer.go:183-186).
func (c *Context) RenderTiles(tiles []uint32) error {
	if len(tiles) == 0 {
		return nil
	}
	if r := C.izpi_render_tiles(c.h, C.int32_t(len(tiles)/4), (*C.uint32_t)(&tiles[0]), nil); r != 0 {
		return lastError("izpi_render_tiles", r)
	}
	return nil
}

// RenderTilesShared is RenderTiles over a tile list shared with other contexts: cursor (C memory, e.g. C.calloc(1, 8),
// initially 0) is the index of the next unclaimed tile and every sharer claims runs of tiles from it -- the work-unit
// channel of renderer.go:126-147.  sharers = number of contexts pulling from the cursor.
func (c *Context) RenderTilesShared(tiles []uint32, cursor *C.uint64_t, sharers int) error {
	if len(tiles) == 0 {
		return nil
	}
	if r := C.izpi_render_tiles_shared(c.h, C.int32_t(len(tiles)/4), (*C.uint32_t)(&tiles[0]), cursor, C.int32_t(sharers), nil); r != 0 {
		return lastError("izpi_render_tiles_shared", r)
	}
	return nil
}

// RenderFinish returns the canvas (Float64NRGBA.Pix layout) and the ray total (renderer.go:213-221).
func (c *Context) RenderFinish(pix []float64) (uint64, error) {
	var rays C.uint64_t
	if r := C.izpi_render_finish(c.h, (*C.double)(&pix[0]), &rays); r != 0 {
		return 0, lastError("izpi_render_finish", r)
	}
	return uint64(rays), nil
}

// RenderTileRows is worker.RenderTile's body (internal/worker/render.go:17-75): it renders the tile and returns the rows
// the worker streams back; row r is image row y0+r and becomes RenderTileResponse{Width: x1-x0+1, Height: 1, PosX: x0,
// PosY: y0+r, Pixels: rows[r*stride:(r+1)*stride]} with stride = stripHeight*4*(x1-x0+1).
func (c *Context) RenderTileRows(stripHeight, x0, y0, x1, y1 uint32) ([]float64, error) {
	stride := int(stripHeight) * 4 * int(x1-x0+1)
	rows := make([]float64, stride*int(y1-y0+1))
	if rc := C.izpi_render_tile_rows(c.h, C.uint32_t(stripHeight), C.uint32_t(x0), C.uint32_t(y0), C.uint32_t(x1), C.uint32_t(y1),
		(*C.double)(ptr(rows))); rc != 0 {
		return nil, lastError("izpi_render_tile_rows", rc)
	}
	return rows, nil
}

// BuildBVH4 is the optional device-side replacement of hitable.NewBVH4 (internal/hitable/bvh4.go:517): boxes holds
// min.xyz max.xyz of every hitable's BoundingBox(time0, time1).  It returns BVH4.Nodes in the reference's 128-byte layout
// and the permutation with Primitives[i] = hitables[perm[i]]; package hitable wraps them into a *BVH4.
func (c *Context) BuildBVH4(boxes []float64) ([]C.izpi_bvh4_node, []int32, error) {
	n := len(boxes) / 6
	var nn C.int32_t
	if rc := C.izpi_bvh4_build(c.h, C.int32_t(n), (*C.double)(ptr(boxes)), &nn); rc != 0 {
		return nil, nil, lastError("izpi_bvh4_build", rc)
	}
	nodes := make([]C.izpi_bvh4_node, int(nn))
	perm := make([]int32, n)
	if rc := C.izpi_bvh4_build_fetch(c.h, ptr(nodes), (*C.int32_t)(ptr(perm))); rc != 0 {
		return nil, nil, lastError("izpi_bvh4_build_fetch", rc)
	}
	return nodes, perm, nil
}

// ApplyDisplacement is displacement.ApplyDisplacementMap (internal/displacement/displacement.go:145) on the device.
// tris15: v0 v1 v2 u0 v0 u1 v1 u2 v2 per triangle; pix: the map as W*H*4 float64 RGBA (height in the blue channel).
func (c *Context) ApplyDisplacement(tris15 []float64, materials []int32, w, h int, pix []float64, min, max float64) ([]float64, []int32, error) {
	var nOut C.int64_t
	if rc := C.izpi_displace(c.h, C.int64_t(len(materials)), (*C.double)(ptr(tris15)), (*C.int32_t)(ptr(materials)), C.int32_t(w), C.int32_t(h),
		(*C.double)(ptr(pix)), C.double(min), C.double(max), 1, &nOut); rc != 0 {
		return nil, nil, lastError("izpi_displace", rc)
	}
	out := make([]float64, 15*int(nOut))
	om := make([]int32, int(nOut))
	if rc := C.izpi_displace_fetch(c.h, (*C.double)(ptr(out)), (*C.int32_t)(ptr(om))); rc != 0 {
		return nil, nil, lastError("izpi_displace_fetch", rc)
	}
	return out, om, nil
}
