function [assignments, distances, centers, info] = skm_lloyd_loop(X, centers, gammaDist, gammaUpdate, varargin)
% SKM_LLOYD_LOOP  The Lloyd iterations of kmeans_sparsified.m:417-486 on the GPU, X resident across iterations.
%
%   [assignments, distances, centers, info] = skm_lloyd_loop(X, centers, gammaDist, gammaUpdate, ...
%                 'MaxIter', 100, 'Tol', 1e-6, 'MLcorrection', true, 'EmptyAction', 'singleton')
%
%   X            sparse p x n matrix, points are COLUMNS (the sparsified matrix after
%                randsample_fixedNumberEntries, kmeans_sparsified.m:334)
%   centers      full p x K start centres (already mixed)
%   gammaDist    SparsityLevel passed to findClusterAssignments (centres are divided by it,
%                private/findClusterAssignments.m:78); [] = no division
%   gammaUpdate  SparsityLevel of the maximum-likelihood centre update (kmeans_sparsified.m:448)
%
% Returns what the reference's loop leaves behind: 1-based assignments (1 x n), Euclidean distances
% (1 x n), the final centres, and info.iterations / info.stoppingDiff / info.objective.
%
% This is the patch of INTEGRATION.md written out as a function, so kmeans_sparsified.m can replace its
% while-loop (:417-486) by one call when the centres are dense.  Sparse centres (the first iterations
% after k-means++ when denseCenters=false, private/findClusterAssignments.m:63-75) stay on the reference
% path or use the Python driver, which routes them through skm_lloyd_assign_sparse.
%
% NOTE: not executed in the development image (no MATLAB); the gateway it calls is tested through the
% stub mex.h (tests/test_mex_shims.py::test_lloyd_gateway_iterates_like_the_reference_loop).

opt = struct('MaxIter', 100, 'Tol', 1e-6, 'MLcorrection', true, 'EmptyAction', 'singleton');
for i = 1:2:numel(varargin)
    if ~isfield(opt, varargin{i}), error('skm_b200:badOption', 'unknown option %s', varargin{i}); end
    opt.(varargin{i}) = varargin{i+1};
end
if ~issparse(X), error('skm_b200:needSparse', 'X must be the sparsified (sparse) matrix'); end
K = size(centers, 2);
h = skm_lloyd_mex('upload', X, K);
cleanup = onCleanup(@() skm_lloyd_mex('free', h));

info = struct('iterations', 0, 'stoppingDiff', NaN, 'objective', NaN);
assignments = []; distances = [];
for its = 1:opt.MaxIter
    centersOld = centers;
    [assignments, distances, centers, dff, sumsq, counts] = ...
        skm_lloyd_mex('iterate', h, full(centers), gammaDist, gammaUpdate, opt.MLcorrection);
    empty = find(counts == 0);
    if ~isempty(empty)                                      % kmeans_sparsified.m:432-445
        warning('kmeans_sparsified:emptyCluster', 'cluster has lost all its members');
        switch lower(opt.EmptyAction)
            case 'singleton'
                [~, iMax] = max(distances);
                for ki = empty(:)'
                    centers(:, ki) = full(X(:, iMax));
                end
            case 'error'
                error('kmeans_sparsified:emptyCluster', 'One cluster lost all its members');
            case 'drop'
                keep = setdiff(1:K, empty);
                centers = centers(:, keep); centersOld = centersOld(:, keep);
                K = numel(keep);
                skm_lloyd_mex('free', h);
                h = skm_lloyd_mex('upload', X, K);          % a state for the smaller K
                assignments = [];
            otherwise
                error('skm_b200:badOption', 'EmptyAction must be singleton, error or drop');
        end
        dff = norm(centersOld - centers, 'fro');            % :470
    end
    info.iterations = its; info.stoppingDiff = dff; info.objective = sqrt(sumsq);   % :470-471
    if dff < opt.Tol, break; end                            % :476-478
    if any(isnan(centers(:))), error('kmeans_sparsified:nan', 'Found NaN in centers'); end
end
if isempty(assignments)                                     % after a 'drop' in the last iteration
    [assignments, distances] = skm_lloyd_mex('assign', h, full(centers), gammaDist);
end
end
