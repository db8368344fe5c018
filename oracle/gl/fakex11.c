/*
 * fakex11.c -- TEST INFRASTRUCTURE (part of oracle/; never linked or loaded by the product).
 *
 * A display-less stand-in for libX11.so.6 / libXext.so.6, just large enough for the Mesa "xlib" software
 * libGL that ships inside the Nsight Compute tree of this image (Mesa 18.1.9, gallium llvmpipe, GLX
 * emulated on Xlib) to create a pbuffer + an OpenGL 3.3 core context without an X server.  With it the
 * reference's own GLSL (pyvr/shaders/volume.{vert,frag}.glsl) runs on Mesa llvmpipe on the host cores,
 * which is the CPU baseline BASELINE.json names and the run that pins oracle/pyvr_oracle.c
 * (oracle/gl/glx_reference.py drives it; it mirrors pyvr/moderngl_renderer/manager.py call for call).
 *
 * Only the 27 Xlib entry points that libGL imports are provided.  All rendering goes to an application
 * FBO (manager.py:29-31), so nothing is ever presented: drawing calls are no-ops, the "server" has one
 * screen with one 24-bit TrueColor visual, and pixmaps are ids with a remembered size.
 * Structure layouts follow <X11/Xlib.h> / <X11/Xutil.h> of X11R7 on LP64 (no X headers in this image).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* FAKEX11_TRACE=1 logs every entry point to stderr (debugging aid) */
static int g_trace = -1;
#define TRACE() do { if (g_trace < 0) g_trace = getenv("FAKEX11_TRACE") != 0; if (g_trace) fprintf(stderr, "[fakex11] %s\n", __func__); } while (0)

typedef unsigned long XID;
typedef XID Window, Drawable, Pixmap, Colormap, VisualID;
typedef char *XPointer;
typedef int Bool;
typedef int Status;
typedef struct _XDisplay Display;
typedef struct _XGC *GC;

typedef struct _XExtData XExtData;

typedef struct {
    XExtData *ext_data;
    VisualID visualid;
    int c_class;
    unsigned long red_mask, green_mask, blue_mask;
    int bits_per_rgb;
    int map_entries;
} Visual;

typedef struct {
    int depth;
    int nvisuals;
    Visual *visuals;
} Depth;

typedef struct {
    XExtData *ext_data;
    Display *display;
    Window root;
    int width, height;
    int mwidth, mheight;
    int ndepths;
    Depth *depths;
    int root_depth;
    Visual *root_visual;
    GC default_gc;
    Colormap cmap;
    unsigned long white_pixel;
    unsigned long black_pixel;
    int max_maps, min_maps;
    int backing_store;
    Bool save_unders;
    long root_input_mask;
} Screen;

typedef struct {
    XExtData *ext_data;
    int depth;
    int bits_per_pixel;
    int scanline_pad;
} ScreenFormat;

typedef struct {
    int extension;
    int major_opcode;
    int first_event;
    int first_error;
} XExtCodes;

/* struct _XExten of Xlibint.h: Mesa's GLX emulation walks dpy->ext_procs, expects XAddExtension to push a
 * new record at its head, and then fills in `name` and `close_display` itself (glx_api.c, register_with_display) */
typedef struct _XExten {
    struct _XExten *next;
    XExtCodes codes;
    void *create_GC, *copy_GC, *flush_GC, *free_GC, *create_Font, *free_Font;
    void *close_display;
    void *error, *error_string;
    char *name;
    void *error_values, *before_flush;
    struct _XExten *next_flush;
} _XExtension;

/* struct _XDisplay as Xlibint.h lays it out, up to ext_procs (the Xlib macros read the public prefix:
 * screens, nscreens, default_screen; Mesa also reads ext_procs) */
struct _XDisplay {
    XExtData *ext_data;
    void *free_funcs;
    int fd;
    int conn_checker;
    int proto_major_version;
    int proto_minor_version;
    char *vendor;
    XID resource_base;
    XID resource_mask;
    XID resource_id;
    int resource_shift;
    XID (*resource_alloc)(Display *);
    int byte_order;
    int bitmap_unit;
    int bitmap_pad;
    int bitmap_bit_order;
    int nformats;
    ScreenFormat *pixmap_format;
    int vnumber;
    int release;
    void *head, *tail;
    int qlen;
    unsigned long last_request_read;
    unsigned long request;
    char *last_req;
    char *buffer;
    char *bufptr;
    char *bufmax;
    unsigned max_request_size;
    void *db;
    int (*synchandler)(Display *);
    char *display_name;
    int default_screen;
    int nscreens;
    Screen *screens;
    unsigned long motion_buffer;
    volatile unsigned long flags;
    int min_keycode;
    int max_keycode;
    void *keysyms;
    void *modifiermap;
    int keysyms_per_keycode;
    char *xdefaults;
    char *scratch_buffer;
    unsigned long scratch_length;
    int ext_number;
    _XExtension *ext_procs;
    char tail_private[4096]; /* the rest of Xlib's private state: zeroed, never interpreted here */
};

typedef struct {
    Visual *visual;
    VisualID visualid;
    int screen;
    int depth;
    int c_class;
    unsigned long red_mask, green_mask, blue_mask;
    int colormap_size;
    int bits_per_rgb;
} XVisualInfo;

typedef struct _XImage {
    int width, height;
    int xoffset;
    int format;
    char *data;
    int byte_order;
    int bitmap_unit;
    int bitmap_bit_order;
    int bitmap_pad;
    int depth;
    int bytes_per_line;
    int bits_per_pixel;
    unsigned long red_mask, green_mask, blue_mask;
    XPointer obdata;
    struct funcs {
        struct _XImage *(*create_image)(Display *, Visual *, unsigned, int, int, char *, unsigned, unsigned, int, int);
        int (*destroy_image)(struct _XImage *);
        unsigned long (*get_pixel)(struct _XImage *, int, int);
        int (*put_pixel)(struct _XImage *, int, int, unsigned long);
        struct _XImage *(*sub_image)(struct _XImage *, int, int, unsigned, unsigned);
        int (*add_pixel)(struct _XImage *, long);
    } f;
} XImage;

typedef struct {
    int x, y;
    int width, height;
    int border_width;
    int depth;
    Visual *visual;
    Window root;
    int c_class;
    int bit_gravity;
    int win_gravity;
    int backing_store;
    unsigned long backing_planes;
    unsigned long backing_pixel;
    Bool save_under;
    Colormap colormap;
    Bool map_installed;
    int map_state;
    long all_event_masks;
    long your_event_mask;
    long do_not_propagate_mask;
    Bool override_redirect;
    Screen *screen;
} XWindowAttributes;

#define TrueColor 4
#define ZPixmap 2

static Visual g_visual;
static Depth g_depth;
static Screen g_screen;
static ScreenFormat g_format;
static Display g_display;
static int g_ready = 0;

#define MAX_DRAWABLES 256
static struct { XID id; unsigned w, h; } g_drawables[MAX_DRAWABLES];
static int g_n_drawables = 0;
static XID g_next_id = 0x400001;

/* Xlib's global lock hooks (LockDisplay / _XLockMutex macros test the function pointer first) */
void (*_XLockMutex_fn)(void *) = 0;
void (*_XUnlockMutex_fn)(void *) = 0;
void *_Xglobal_lock = 0;

static void init_display(void) {
    if (g_ready) return;
    memset(&g_display, 0, sizeof g_display);
    g_visual.visualid = 0x21;
    g_visual.c_class = TrueColor;
    g_visual.red_mask = 0xff0000;
    g_visual.green_mask = 0x00ff00;
    g_visual.blue_mask = 0x0000ff;
    g_visual.bits_per_rgb = 8;
    g_visual.map_entries = 256;
    g_depth.depth = 24;
    g_depth.nvisuals = 1;
    g_depth.visuals = &g_visual;
    g_screen.display = &g_display;
    g_screen.root = 0x100;
    g_screen.width = 1920;
    g_screen.height = 1080;
    g_screen.mwidth = 508;
    g_screen.mheight = 286;
    g_screen.ndepths = 1;
    g_screen.depths = &g_depth;
    g_screen.root_depth = 24;
    g_screen.root_visual = &g_visual;
    g_screen.cmap = 0x20;
    g_screen.white_pixel = 0xffffff;
    g_format.depth = 24;
    g_format.bits_per_pixel = 32;
    g_format.scanline_pad = 32;
    g_display.fd = -1;
    g_display.proto_major_version = 11;
    g_display.vendor = (char *)"pyvr_b200 oracle (no server)";
    g_display.byte_order = 0;
    g_display.bitmap_unit = 32;
    g_display.bitmap_pad = 32;
    g_display.bitmap_bit_order = 0;
    g_display.nformats = 1;
    g_display.pixmap_format = &g_format;
    g_display.release = 12101000;
    g_display.max_request_size = 65535;
    g_display.display_name = (char *)":fake";
    g_display.default_screen = 0;
    g_display.nscreens = 1;
    g_display.screens = &g_screen;
    g_ready = 1;
}

Display *XOpenDisplay(const char *name) { TRACE(); (void)name; init_display(); return &g_display; }
int XCloseDisplay(Display *d) { TRACE(); (void)d; return 0; }

XExtCodes *XAddExtension(Display *d) { TRACE();
    _XExtension *ext = (_XExtension *)calloc(1, sizeof(_XExtension) + 64);
    if (!ext) return 0;
    ext->codes.extension = d->ext_number++;
    ext->next = d->ext_procs;
    d->ext_procs = ext;
    return &ext->codes;
}
/* close-display hook registration: accepted and never called (the display is never closed) */
void *XESetCloseDisplay(Display *d, int ext, void *proc) { TRACE(); (void)d; (void)ext; (void)proc; return 0; }

Colormap XCreateColormap(Display *d, Window w, Visual *v, int alloc) { TRACE(); (void)d; (void)w; (void)v; (void)alloc; return 0x21; }
GC XCreateGC(Display *d, Drawable dr, unsigned long mask, void *values) { TRACE(); (void)d; (void)dr; (void)mask; (void)values; return (GC)calloc(1, 256); }
int XFreeGC(Display *d, GC gc) { TRACE(); (void)d; free(gc); return 1; }

static int image_destroy(XImage *img) {
    if (img) { free(img->data); free(img); }
    return 1;
}
static unsigned long image_get_pixel(XImage *img, int x, int y) {
    if (!img || !img->data) return 0;
    const unsigned char *p = (const unsigned char *)img->data + (size_t)y * img->bytes_per_line + (size_t)x * (img->bits_per_pixel / 8);
    unsigned long v = 0;
    memcpy(&v, p, img->bits_per_pixel / 8);
    return v;
}
static int image_put_pixel(XImage *img, int x, int y, unsigned long v) {
    if (!img || !img->data) return 0;
    unsigned char *p = (unsigned char *)img->data + (size_t)y * img->bytes_per_line + (size_t)x * (img->bits_per_pixel / 8);
    memcpy(p, &v, img->bits_per_pixel / 8);
    return 1;
}
static XImage *image_sub(XImage *img, int x, int y, unsigned w, unsigned h) { (void)img; (void)x; (void)y; (void)w; (void)h; return 0; }
static int image_add_pixel(XImage *img, long v) { (void)img; (void)v; return 0; }

XImage *XCreateImage(Display *d, Visual *visual, unsigned depth, int format, int offset, char *data,
                     unsigned width, unsigned height, int bitmap_pad, int bytes_per_line) { TRACE();
    (void)d;
    XImage *img = (XImage *)calloc(1, sizeof(XImage));
    if (!img) return 0;
    img->width = (int)width;
    img->height = (int)height;
    img->xoffset = offset;
    img->format = format;
    img->data = data;
    img->byte_order = 0;
    img->bitmap_unit = 32;
    img->bitmap_bit_order = 0;
    img->bitmap_pad = bitmap_pad ? bitmap_pad : 32;
    img->depth = (int)depth;
    img->bits_per_pixel = depth <= 8 ? 8 : depth <= 16 ? 16 : 32;
    if (bytes_per_line == 0) {
        const unsigned pad = (unsigned)img->bitmap_pad;
        bytes_per_line = (int)(((width * (unsigned)img->bits_per_pixel + pad - 1) / pad) * pad / 8);
    }
    img->bytes_per_line = bytes_per_line;
    if (visual) { img->red_mask = visual->red_mask; img->green_mask = visual->green_mask; img->blue_mask = visual->blue_mask; }
    img->f.create_image = 0;
    img->f.destroy_image = image_destroy;
    img->f.get_pixel = image_get_pixel;
    img->f.put_pixel = image_put_pixel;
    img->f.sub_image = image_sub;
    img->f.add_pixel = image_add_pixel;
    return img;
}

Pixmap XCreatePixmap(Display *d, Drawable dr, unsigned w, unsigned h, unsigned depth) { TRACE();
    (void)d; (void)dr; (void)depth;
    const XID id = g_next_id++;
    if (g_n_drawables < MAX_DRAWABLES) {
        g_drawables[g_n_drawables].id = id;
        g_drawables[g_n_drawables].w = w;
        g_drawables[g_n_drawables].h = h;
        ++g_n_drawables;
    }
    return id;
}
int XFreePixmap(Display *d, Pixmap p) { TRACE();
    (void)d;
    for (int i = 0; i < g_n_drawables; ++i)
        if (g_drawables[i].id == p) { g_drawables[i] = g_drawables[--g_n_drawables]; break; }
    return 1;
}

int XDrawString16(Display *d, Drawable dr, GC gc, int x, int y, const void *s, int n) { TRACE(); (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)s; (void)n; return 0; }
int XFillRectangle(Display *d, Drawable dr, GC gc, int x, int y, unsigned w, unsigned h) { TRACE(); (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)w; (void)h; return 1; }
int XFlush(Display *d) { TRACE(); (void)d; return 1; }
int XFree(void *p) { TRACE(); free(p); return 1; }
int XFreeFontInfo(char **names, void *info, int n) { TRACE(); (void)names; (void)info; (void)n; return 1; }

Status XGetGeometry(Display *d, Drawable dr, Window *root, int *x, int *y, unsigned *w, unsigned *h,
                    unsigned *border, unsigned *depth) { TRACE();
    (void)d;
    unsigned ww = 1, hh = 1;
    for (int i = 0; i < g_n_drawables; ++i)
        if (g_drawables[i].id == dr) { ww = g_drawables[i].w; hh = g_drawables[i].h; }
    if (root) *root = g_screen.root;
    if (x) *x = 0;
    if (y) *y = 0;
    if (w) *w = ww;
    if (h) *h = hh;
    if (border) *border = 0;
    if (depth) *depth = 24;
    return 1;
}

XImage *XGetImage(Display *d, Drawable dr, int x, int y, unsigned w, unsigned h, unsigned long mask, int format) { TRACE();
    (void)d; (void)dr; (void)x; (void)y; (void)w; (void)h; (void)mask; (void)format;
    return 0;
}

XVisualInfo *XGetVisualInfo(Display *d, long mask, XVisualInfo *t, int *n) { TRACE();
    (void)d;
    init_display();
    int ok = 1;
    if (t) {
        if ((mask & 0x1) && t->visualid != g_visual.visualid) ok = 0;
        if ((mask & 0x2) && t->screen != 0) ok = 0;
        if ((mask & 0x4) && t->depth != 24) ok = 0;
        if ((mask & 0x8) && t->c_class != TrueColor) ok = 0;
        if ((mask & 0x10) && t->red_mask != g_visual.red_mask) ok = 0;
        if ((mask & 0x20) && t->green_mask != g_visual.green_mask) ok = 0;
        if ((mask & 0x40) && t->blue_mask != g_visual.blue_mask) ok = 0;
        if ((mask & 0x80) && t->colormap_size != 256) ok = 0;
        if ((mask & 0x100) && t->bits_per_rgb != 8) ok = 0;
    }
    if (!ok) { if (n) *n = 0; return 0; }
    XVisualInfo *v = (XVisualInfo *)calloc(1, sizeof(XVisualInfo));
    v->visual = &g_visual;
    v->visualid = g_visual.visualid;
    v->screen = 0;
    v->depth = 24;
    v->c_class = TrueColor;
    v->red_mask = g_visual.red_mask;
    v->green_mask = g_visual.green_mask;
    v->blue_mask = g_visual.blue_mask;
    v->colormap_size = 256;
    v->bits_per_rgb = 8;
    if (n) *n = 1;
    return v;
}

Status XGetWindowAttributes(Display *d, Window w, XWindowAttributes *a) { TRACE();
    (void)d;
    if (!a) return 0;
    memset(a, 0, sizeof *a);
    unsigned ww = 1, hh = 1;
    for (int i = 0; i < g_n_drawables; ++i)
        if (g_drawables[i].id == w) { ww = g_drawables[i].w; hh = g_drawables[i].h; }
    a->width = (int)ww;
    a->height = (int)hh;
    a->depth = 24;
    a->visual = &g_visual;
    a->root = g_screen.root;
    a->c_class = 1;
    a->colormap = g_screen.cmap;
    a->map_state = 2;
    a->screen = &g_screen;
    return 1;
}

int XPutImage(Display *d, Drawable dr, GC gc, XImage *img, int sx, int sy, int dx, int dy, unsigned w, unsigned h) { TRACE();
    (void)d; (void)dr; (void)gc; (void)img; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h;
    return 0;
}
Bool XQueryExtension(Display *d, const char *name, int *major, int *event, int *error) { TRACE();
    (void)d; (void)name;
    if (major) *major = 0;
    if (event) *event = 0;
    if (error) *error = 0;
    return 0;
}
void *XQueryFont(Display *d, XID id) { TRACE(); (void)d; (void)id; return 0; }
typedef int (*XErrorHandler)(Display *, void *);
XErrorHandler XSetErrorHandler(XErrorHandler h) { TRACE(); static XErrorHandler cur = 0; XErrorHandler prev = cur; cur = h; return prev; }
int XSetForeground(Display *d, GC gc, unsigned long fg) { TRACE(); (void)d; (void)gc; (void)fg; return 1; }
int XSetFunction(Display *d, GC gc, int fn) { TRACE(); (void)d; (void)gc; (void)fn; return 1; }
int XSync(Display *d, Bool discard) { TRACE(); (void)d; (void)discard; return 1; }
typedef int (*XAfterFunction)(Display *);
XAfterFunction XSynchronize(Display *d, Bool onoff) { TRACE(); (void)d; (void)onoff; return 0; }

/* MIT-SHM (libXext): reported absent by XQueryExtension; these fail if called anyway */
Bool XShmQueryExtension(Display *d) { TRACE(); (void)d; return 0; }
Status XShmAttach(Display *d, void *info) { TRACE(); (void)d; (void)info; return 0; }
Status XShmDetach(Display *d, void *info) { TRACE(); (void)d; (void)info; return 0; }
XImage *XShmCreateImage(Display *d, Visual *v, unsigned depth, int format, char *data, void *info, unsigned w, unsigned h) { TRACE();
    (void)d; (void)v; (void)depth; (void)format; (void)data; (void)info; (void)w; (void)h;
    return 0;
}
Status XShmPutImage(Display *d, Drawable dr, GC gc, XImage *img, int sx, int sy, int dx, int dy, unsigned w, unsigned h, Bool ev) { TRACE();
    (void)d; (void)dr; (void)gc; (void)img; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; (void)ev;
    return 0;
}
