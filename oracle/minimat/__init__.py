"""minimat -- a small interpreter for the subset of MATLAB that ndefilippis/SWRaytracing's hot path is written in.

TEST INFRASTRUCTURE ONLY (like everything under ``oracle/``): only ``tests/`` and the golden-vector generators under
``tests/golden/`` may import it; the product (``swraytracing_b200``) never does.

Why it exists.  The reference is MATLAB and neither MATLAB nor GNU Octave is installed in the build image or on the GPU box, so
the reference's own code could not be executed and the CPU oracle (``oracle/swrt_oracle.py`` / ``.c``, a hand restatement) had
nothing of the reference's to be checked against for the packet arithmetic.  ``minimat`` closes that gap: it parses and
executes the UNMODIFIED ``.m`` files where they lie under ``/root/reference`` -- ``ray_trace_sw/interpolate.m``,
``qg_flow_ray_trace/interpolate_U.m``, ``SpectralScheme.m`` / ``RaytracingScheme.m`` (classdef + inheritance),
``ode_symplectic.m`` (nested functions, function handles, closures), ``ray_trace_sw/cg_sw.m``, ``step_packet.m``,
``step_packet_xka.m``, ``rsw/g2k.m`` / ``k2g.m`` / ``fulspec.m``, ``qg_flow_ray_trace/read_field.m`` / ``write_field.m`` --
driven by the committed recipe ``tests/golden/make_octave_goldens.m`` (also unmodified; it runs under real MATLAB / Octave
too).  ``tests/golden/run_reference_recipe.py`` is the one command; its output is ``tests/golden/octave_out/*.bin``.

What makes its output trustworthy.  The interpreter is generic MATLAB semantics (column-major arrays, 1-based / logical /
``end`` indexing, growth on assignment, implicit expansion, value semantics, ``nargin`` / ``nargout``, nested-function variable
sharing, method dispatch, path precedence, ``fread`` / ``fwrite``), not a restatement of the path, and it is itself pinned
to numbers that REAL MATLAB produced and the reference still holds (``tests/test_minimat.py``):

* running the unmodified ``qg_flow_ray_trace/qgsw_raytrace.m`` with the arguments of the shipped SLURM runs prints the header
  lines MATLAB R2020b printed into ``run.log`` / ``analysis/job-*/run-*/run.log`` character for character -- every line that
  depends on arithmetic (``rng(146)`` -> ``rand`` -> ``initial_q`` with its always-true chained comparison -> ``g2k`` ->
  ``grid_U`` -> 6 x ``k2g`` -> ``U0``, ``Fr``, ``dt``; 'Simulation time' is skipped, the logged runs defined T differently) --
  and the ``pv_time`` frames its solver loop writes land within 2 ulps of the stream the reference's own run stored;
* running the unmodified ``rsw/swk.m`` -- a 359-line pseudo-spectral rotating-shallow-water solver -- from the initial condition
  that ``rsw/matlab.mat`` (MATLAB's own workspace dump at step 300 of that run) still holds, for the same 300 steps, reproduces
  the clock, time step, viscosity and counters MATLAB held EXACTLY, the spectral state to 1e-15 and the energy series to 1e-13;
* running the unmodified ``rsw/k2g.m`` / ``fulspec.m`` on the spectral state of ``rsw/matlab.mat`` reproduces the grid fields
  MATLAB stored in the same workspace to 1e-15.

Floating point: every arithmetic operation is IEEE double through numpy / libm, like MATLAB's; library calls whose last bits
may differ from MATLAB's (FFT, sum order of long vectors, ``pow``) are listed in ``builtins.py``.  The hot-path functions
use none of them inside the per-packet arithmetic, and the tests hold results to 1e-12 / 1e-9, not to the bit.

What it is not.  A subset, grown until the reference ran: double / complex / logical / char arrays, cells, structs and 2-D struct
arrays, function handles, value classes (``classdef`` with inheritance; ``handle`` classes run with value semantics, which is all
the shims need).  No integer or single classes, no sparse matrices, no ``try`` beyond catching the interpreter's own errors, no
graphics (plotting calls are accepted and ignored), no ODE suite (``ode23`` is a MATLAB builtin, not reference code: where a
driver needs it the test supplies the restated controller and says so), no ``eval`` / ``evalin`` / ``inputname``.  ``end`` inside
``x(...)`` of a function CALL is not supported (only inside subscripts of variables), and command syntax follows MATLAB's rule
for names that are not variables.  Anything outside the subset raises ``MatlabError`` / ``ParseError`` -- it never guesses.

Layout: ``lexer.py`` (tokens, transpose-vs-quote, command syntax), ``parser.py`` (statements, functions with and without
``end``, nested functions, classdef), ``values.py`` (array semantics), ``interp.py`` (evaluator), ``builtins.py`` (library).
"""
from .interp import Interp, Frame, from_py, to_py          # noqa: F401
from .values import MatlabError, MStruct, MCell, MObject, FuncHandle   # noqa: F401
