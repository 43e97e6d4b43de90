-- | Drop-in for "Data.Text.AhoCorasick.Searcher" (reference: src/Data/Text/AhoCorasick/Searcher.hs:14-27).
-- NOT COMPILED HERE -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Searcher
  ( Searcher, build, buildWithValues, buildNeedleIdSearcher, containsAny, containsAll
  , needles, numNeedles, automaton, caseSensitivity, setCaseSensitivity, mapSearcher
  ) where

import Data.Hashable (Hashable)
import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (Text)
import Foreign.ForeignPtr (withForeignPtr)
import Foreign.Marshal (alloca)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Text.AhoCorasick.Automaton as Aho
import Data.Text.AhoCorasick.FFI

data Searcher v = Searcher                         -- Searcher.hs:61-66
  { searcherCaseSensitive :: CaseSensitivity
  , searcherNeedles :: [(Text, v)]
  , searcherNumNeedles :: Int
  , searcherAutomaton :: Aho.AcMachine v
  }

build :: CaseSensitivity -> [Text] -> Searcher ()                                   -- :110-111
build case_ = buildWithValues case_ . fmap (\n -> (n, ()))

buildWithValues :: Hashable v => CaseSensitivity -> [(Text, v)] -> Searcher v      -- :115-118 (IgnoreCase: the caller lower-cases, :107-109)
buildWithValues case_ ns = Searcher case_ ns (length ns) (Aho.build ns)

buildNeedleIdSearcher :: CaseSensitivity -> [Text] -> Searcher Int                 -- :167-169
buildNeedleIdSearcher case_ ns = buildWithValues case_ (zip ns [0 ..])

needles :: Searcher v -> [(Text, v)]
needles = searcherNeedles
numNeedles :: Searcher v -> Int
numNeedles = searcherNumNeedles
automaton :: Searcher v -> Aho.AcMachine v
automaton = searcherAutomaton
caseSensitivity :: Searcher v -> CaseSensitivity
caseSensitivity = searcherCaseSensitive

-- | :142-145: flips the flag; needles and automaton are shared -- the machine serves both modes, nothing is rebuilt.
setCaseSensitivity :: CaseSensitivity -> Searcher v -> Searcher v
setCaseSensitivity case_ s = s { searcherCaseSensitive = case_ }

mapSearcher :: (a -> b) -> Searcher a -> Searcher b                                 -- :121-125 (payloads live on the host)
mapSearcher f s = s { searcherNeedles = fmap (fmap f) (searcherNeedles s), searcherAutomaton = fmap f (searcherAutomaton s) }

-- | `containsAny` (:156-164): one am_contains_any call.  The kernel sets a flag at the first match, the other CTAs
-- stop at their next tile and the upload of the haystack stops with them: the fold's `Done True`.
containsAny :: Searcher () -> Text -> Bool
containsAny s text = unsafePerformIO $ withForeignPtr (Aho.machineHandle (automaton s)) $ \h ->
  withSlice text $ \hay -> alloca $ \out -> do
    _ <- amCall "am_contains_any" [] $ c_am_contains_any h (caseToC (caseSensitivity s)) hay out
    (/= 0) <$> peek out
{-# NOINLINE containsAny #-}

-- | `containsAll` (:173-187), for searchers from `buildNeedleIdSearcher`: a bit per needle on the device.
containsAll :: Searcher Int -> Text -> Bool
containsAll s text = unsafePerformIO $ withForeignPtr (Aho.machineHandle (automaton s)) $ \h ->
  withSlice text $ \hay -> alloca $ \out -> do
    _ <- amCall "am_contains_all" [] $ c_am_contains_all h (caseToC (caseSensitivity s)) hay out
    (/= 0) <$> peek out
{-# NOINLINE containsAll #-}
