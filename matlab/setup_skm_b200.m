function setup_skm_b200(repoRoot)
% SETUP_SKM_B200  Build the MEX gateways of libskm_b200 (the B200 engine) and put them on the path.
%
%   setup_skm_b200            uses the directory above this file as the repository root
%   setup_skm_b200(repoRoot)
%
% Counterpart of the reference's setup_kmeans.m:19-57 (which compiles SparseMatrixMinusCluster.c,
% SparseMatrixInnerProduct.c, SparseMatrixColumnNormSq.c, hadamard.c and hadamard_pthreads.c with
% `mex -largeArrayDims`).  The five gateways keep their names and usage, so private/findClusterAssignments.m
% and kmeans_sparsified.m of the reference call them unchanged; skm_lloyd_mex and skm_second_pass_mex are
% the handle-based additions described in INTEGRATION.md.
%
% libskm_b200.so must have been built first:  python -m sparsifiedkmeans_b200.build
%
% NOTE: MATLAB is not installed in the image this repository was developed in.  The identical C sources
% are compiled against a stub mex.h and driven through ctypes in tests/test_mex_shims.py; this script is
% the recipe a MATLAB user runs, it has not been executed here.

if nargin < 1 || isempty(repoRoot)
    repoRoot = fileparts(fileparts(mfilename('fullpath')));
end
inc  = ['-I', fullfile(repoRoot, 'include')];
libd = fullfile(repoRoot, 'sparsifiedkmeans_b200', '_lib');
if ~exist(fullfile(libd, 'libskm_b200.so'), 'file')
    error('skm_b200:notBuilt', 'libskm_b200.so not found in %s; run "python -m sparsifiedkmeans_b200.build" first', libd);
end
lib  = {['-L', libd], '-lskm_b200', ['LDFLAGS=$LDFLAGS -Wl,-rpath,', libd]};
src  = fullfile(repoRoot, 'mex');
out  = fullfile(repoRoot, 'matlab', 'mexbin');
if ~exist(out, 'dir'), mkdir(out); end

names = {'SparseMatrixMinusCluster', 'SparseMatrixInnerProduct', 'SparseMatrixColumnNormSq', ...
         'hadamard', 'skm_lloyd_mex', 'skm_second_pass_mex'};
for i = 1:numel(names)
    mex('-largeArrayDims', '-O', inc, fullfile(src, [names{i}, '.c']), lib{:}, '-outdir', out);
end
% the reference also ships hadamard_pthreads; the same gateway serves both names
mex('-largeArrayDims', '-O', inc, fullfile(src, 'hadamard.c'), lib{:}, '-outdir', out, '-output', 'hadamard_pthreads');

addpath(out);              % must come BEFORE the reference's private/ copies on the path
addpath(fullfile(repoRoot, 'matlab'));
fprintf('skm_b200 gateways built in %s\n', out);
end
